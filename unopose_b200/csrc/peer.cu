// Generic peer-memory all-gather (csrc/peer.cuh): every rank stores `bytes` of its own data into slot [rank] of every
// rank's slab and publishes; upk_peer_wait holds the stream until all ranks have published and optionally copies the
// gathered array out.  Used for the per-step result rows of instance sharding (B x 13 floats per rank) in place of an
// NCCL all_gather.
#include <stdint.h>

#include "common.cuh"
#include "launch_count.h"
#include "peer.cuh"
#include "pose_internal.h"

namespace upk {

__global__ void __launch_bounds__(256)
k_peer_all_gather(const uint4* __restrict__ src, size_t n16, const PeerCtx pc, size_t peer_off, size_t slab_bytes,
                  int channel) {
  const unsigned long long e_new = pc.epoch[channel] + 1;
  const size_t o = peer_off + (e_new & 1) * slab_bytes + (size_t)pc.rank * n16 * sizeof(uint4);
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n16; i += (size_t)gridDim.x * 256) {
    const uint4 v = src[i];
    for (int r = 0; r < pc.world; ++r) reinterpret_cast<uint4*>(pc.data[r] + o)[i] = v;
  }
  peer_publish(pc, channel, e_new, gridDim.x);
}

__global__ void __launch_bounds__(256)
k_peer_wait(const PeerCtx pc, size_t peer_off, size_t slab_bytes, int channel, uint4* __restrict__ dst, size_t n16) {
  const unsigned long long e = pc.epoch[channel];
  peer_wait(pc, channel, e);
  if (!dst) return;
  const uint4* s = reinterpret_cast<const uint4*>(pc.data[pc.rank] + peer_off + (e & 1) * slab_bytes);
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n16; i += (size_t)gridDim.x * 256) dst[i] = __ldcg(s + i);
}

}  // namespace upk

using namespace upk;

extern "C" {

int upk_peer_all_gather(const void* src, size_t bytes, const upk_peer_t* peer, size_t data_offset, size_t slab_bytes,
                        int channel, upk_stream_t stream) {
  if (!src || bytes == 0 || (bytes & 15) || !peer_ctx_ok(peer, channel)) return UPK_ERR_INVALID_ARG;
  if (((data_offset | slab_bytes) & 15) || ((uintptr_t)src & 15) || bytes * peer->world > slab_bytes) return UPK_ERR_INVALID_ARG;
  const size_t n16 = bytes / 16;
  const unsigned grid = (unsigned)(n16 / 256 + 1 < 32 ? n16 / 256 + 1 : 32);
  k_peer_all_gather<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)src, n16, make_peer_ctx(peer), data_offset,
                                                            slab_bytes, channel);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

int upk_peer_wait(const upk_peer_t* peer, size_t data_offset, size_t slab_bytes, int channel, void* dst, size_t bytes,
                  upk_stream_t stream) {
  if (!peer_ctx_ok(peer, channel) || ((data_offset | slab_bytes) & 15)) return UPK_ERR_INVALID_ARG;
  if (dst && ((bytes & 15) || ((uintptr_t)dst & 15) || bytes > slab_bytes)) return UPK_ERR_INVALID_ARG;
  const size_t n16 = dst ? bytes / 16 : 0;
  const unsigned grid = (unsigned)(n16 / 256 + 1 < 32 ? n16 / 256 + 1 : 32);
  k_peer_wait<<<grid, 256, 0, (cudaStream_t)stream>>>(make_peer_ctx(peer), data_offset, slab_bytes, channel, (uint4*)dst, n16);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

}  // extern "C"
