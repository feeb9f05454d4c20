// Peer exchange over NVLink / NVSwitch peer memory: the producing kernel of an exchange step stores its
// results straight into EVERY rank's copy of a symmetric buffer (plain st.global to peer-mapped addresses), the last
// CTA of the grid publishes an epoch flag on every peer, and the consuming kernel spins on its LOCAL flags before it
// reads — one kernel does the compute AND the all-gather, no NCCL launch on the critical path (an NCCL all_gather of
// these 100 KB messages costs ~17 us on 2 B200s; the store + flag round trip is a few us).
//
// Protocol (one "channel" per exchange step of a solve):
//   epoch[ch]   this rank's count of completed publications (device memory, advanced by the producer kernel itself, so
//               a captured CUDA graph replays correctly);
//   slab parity data of epoch e live in slab (e & 1): a rank can be at most one publication ahead of a peer that is
//               still reading (its publication e+2 needs the peer's e+1, which the peer issues after it consumed e);
//   flags       flags[r][ch * UPK_MAX_PEERS + src] = latest epoch published by rank `src`, resident on rank r.
// Ordering: every producer thread fences at system scope after its peer stores, the CTA barrier and a device-scope
// atomic counter elect the last CTA, which fences again and releases the flags at system scope; consumers acquire at
// system scope and read the exchanged data past L1 (ld.global.cg).  A bounded spin (about 2 s) raises status[0] instead
// of hanging the GPU when a peer never arrives.
#pragma once
#include <cuda_runtime.h>

#include "../../include/unopose_b200.h"

namespace upk {

struct PeerCtx {
  int world, rank;
  char* data[UPK_MAX_PEERS];
  unsigned long long* flags[UPK_MAX_PEERS];
  unsigned long long* epoch;
  unsigned int* done;
  unsigned int* status;
};

inline PeerCtx make_peer_ctx(const upk_peer_t* p) {
  PeerCtx c;
  c.world = p->world;
  c.rank = p->rank;
  for (int r = 0; r < UPK_MAX_PEERS; ++r) {
    c.data[r] = (char*)p->data[r];
    c.flags[r] = (unsigned long long*)p->flags[r];
  }
  c.epoch = (unsigned long long*)p->epoch;
  c.done = (unsigned int*)p->done;
  c.status = (unsigned int*)p->status;
  return c;
}

inline bool peer_ctx_ok(const upk_peer_t* p, int channel) {
  if (!p || p->world < 1 || p->world > UPK_MAX_PEERS || p->rank < 0 || p->rank >= p->world) return false;
  if (channel < 0 || channel >= UPK_PEER_CHANNELS || !p->epoch || !p->done || !p->status) return false;
  for (int r = 0; r < p->world; ++r)
    if (!p->data[r] || !p->flags[r]) return false;
  return true;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Every thread of every CTA calls this after its peer stores of epoch e_new on `channel`.
__device__ __forceinline__ void peer_publish(const PeerCtx& p, int channel, unsigned long long e_new, unsigned total_ctas) {
  __threadfence_system();
  __syncthreads();
  const unsigned tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  if (tid == 0) {
    const unsigned prev = atomicAdd(&p.done[channel], 1u);
    if (prev == total_ctas - 1) {
      p.done[channel] = 0;              // the counter is ready for the next launch
      __threadfence_system();
      p.epoch[channel] = e_new;
      for (int r = 0; r < p.world; ++r)
        st_release_sys(p.flags[r] + channel * UPK_MAX_PEERS + p.rank, e_new);
    }
  }
}

// First thing a consuming CTA does: wait until every rank has published epoch e on `channel`.
__device__ __forceinline__ void peer_wait(const PeerCtx& p, int channel, unsigned long long e) {
  const unsigned tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  if (tid < (unsigned)p.world && (int)tid != p.rank) {
    const unsigned long long* f = p.flags[p.rank] + channel * UPK_MAX_PEERS + tid;
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < e) {
      if (clock64() - t0 > 4000000000LL) { atomicExch(p.status, 1u); break; }   // ~2 s: report, do not hang
      __nanosleep(64);
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }

}  // namespace upk
