"""Peer-memory exchange between the GPUs of one node (NVLink / NVSwitch), the host side of csrc/peer.cuh.

One SYMMETRIC allocation per rank (torch.distributed._symmetric_memory: CUDA VMM memory that every process of the group
maps) holds the flag array and the exchange slabs; `upk_peer_t` hands the peer-mapped addresses to the kernels, which
store their results into every rank's copy and publish an epoch flag — the fused compute + all-gather form of
 * the two exchange steps of hypothesis sharding (dist.HypothesisShardedCoarse, exchange="p2p"), and
 * the per-step result gather of instance sharding (PeerAllGather)
(SURVEY.md §8e).  Everything has static addresses and the epochs advance on the device, so the exchange is captured
into CUDA graphs together with the kernels around it.  NCCL remains the transport where symmetric memory is not
available (CPU/gloo tests, GPUs without peer access)."""
import ctypes

import torch
import torch.distributed as dist

from . import _lib as L

MAX_PEERS = 8
CHANNELS = 4
_HEADER = 1024          # flag array [CHANNELS][MAX_PEERS] u64 = 256 B, padded


class PeerStruct(ctypes.Structure):
    """struct upk_peer (include/unopose_b200.h)"""
    _fields_ = [("world", ctypes.c_int), ("rank", ctypes.c_int),
                ("data", ctypes.c_void_p * MAX_PEERS), ("flags", ctypes.c_void_p * MAX_PEERS),
                ("epoch", ctypes.c_void_p), ("done", ctypes.c_void_p), ("status", ctypes.c_void_p)]


def available(device, group=None):
    """True when the ranks of `group` (one node) can map each other's memory."""
    try:
        if not (dist.is_available() and dist.is_initialized()) or torch.device(device).type != "cuda":
            return False
        world = dist.get_world_size(group)
        if world < 2 or world > MAX_PEERS or "nccl" not in str(dist.get_backend(group)):
            return False
        import torch.distributed._symmetric_memory as _symm  # noqa: F401

        me = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
        return all(torch.cuda.can_device_access_peer(me, j) for j in range(torch.cuda.device_count()) if j != me)
    except Exception as ex:  # noqa: BLE001
        import os

        if os.environ.get("UPK_PEER_DEBUG"):
            print("peer.available: %r" % (ex,), flush=True)
        return False


class PeerExchange:
    """The symmetric buffer of one rank + the `upk_peer_t` describing all ranks' buffers.  `reserve(nbytes)` carves a
    channel: two slabs of `nbytes` (epoch parity) -> (channel, data_offset, slab_bytes)."""

    def __init__(self, data_bytes, device, group=None):
        import torch.distributed._symmetric_memory as symm

        self.group = group
        self.device = torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > MAX_PEERS:
            raise RuntimeError("peer exchange supports up to %d ranks" % MAX_PEERS)
        self.capacity = (int(data_bytes) + 1023) // 1024 * 1024
        self.buf = symm.empty(_HEADER + self.capacity, dtype=torch.uint8, device=self.device)
        name = (group if group is not None else dist.group.WORLD).group_name
        self.hdl = symm.rendezvous(self.buf, name)
        self.buf.zero_()
        # local counters: epoch [CHANNELS] u64 | done [CHANNELS] u32 | status u32
        self.local = torch.zeros(CHANNELS * 8 + CHANNELS * 4 + 16, dtype=torch.uint8, device=self.device)
        torch.cuda.synchronize(self.device)
        dist.barrier(group)                # nobody publishes before every rank has zeroed its flags
        s = PeerStruct()
        s.world, s.rank = self.world, self.rank
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        for r in range(self.world):
            s.flags[r] = ptrs[r]
            s.data[r] = ptrs[r] + _HEADER
        base = self.local.data_ptr()
        s.epoch, s.done, s.status = base, base + CHANNELS * 8, base + CHANNELS * 8 + CHANNELS * 4
        self.struct = s
        self.ref = ctypes.byref(s)
        self._used = 0
        self._channels = 0

    def reserve(self, nbytes):
        slab = (int(nbytes) + 255) // 256 * 256
        if self._channels >= CHANNELS or self._used + 2 * slab > self.capacity:
            raise RuntimeError("peer exchange: out of channels / buffer space")
        ch, off = self._channels, self._used
        self._channels += 1
        self._used += 2 * slab
        return ch, off, slab

    def timed_out(self):
        """True if any wait gave up (a peer never published) since construction; synchronises."""
        return bool(self.local[CHANNELS * 12:CHANNELS * 12 + 4].view(torch.int32).item())


class PeerAllGather:
    """all_gather of a fixed-size tensor per rank through peer memory: `gather(x)` -> (world, *x.shape) (a static
    output buffer).  One store kernel (into every rank's slab, publishes) + one wait/copy kernel."""

    def __init__(self, shape, dtype, device, group=None, exchange=None):
        self.shape, self.dtype = tuple(shape), dtype
        n = int(torch.tensor(self.shape).prod().item()) * torch.empty((), dtype=dtype).element_size()
        self.bytes = (n + 15) // 16 * 16
        self.nbytes = n
        self.world = dist.get_world_size(group)
        self.px = exchange or PeerExchange(2 * self.bytes * self.world + 1024, device, group)
        self.ch, self.off, self.slab = self.px.reserve(self.bytes * self.world)
        self.stage = torch.zeros(self.bytes, dtype=torch.uint8, device=device)
        self.out = torch.zeros(self.world * self.bytes, dtype=torch.uint8, device=device)
        self.lib = L.load()

    def gather(self, x):
        if tuple(x.shape) != self.shape or x.dtype != self.dtype:
            raise RuntimeError("PeerAllGather: expected %s %s" % (self.shape, self.dtype))
        st = L.stream_ptr(x)
        src = x.contiguous()
        if self.nbytes != self.bytes or src.data_ptr() % 16:
            self.stage[:self.nbytes].copy_(src.view(-1).view(torch.uint8))
            src = self.stage
        with torch.cuda.device(x.device):
            L.check(self.lib.upk_peer_all_gather(src.data_ptr(), self.bytes, self.px.ref, self.off, self.slab, self.ch, st),
                    "peer_all_gather")
            L.check(self.lib.upk_peer_wait(self.px.ref, self.off, self.slab, self.ch, self.out.data_ptr(),
                                           self.world * self.bytes, st), "peer_wait")
        o = self.out.view(self.world, self.bytes)[:, :self.nbytes]
        return o.reshape(-1).view(self.dtype).view((self.world,) + self.shape) if self.nbytes == self.bytes else \
            o.contiguous().view(-1).view(self.dtype).view((self.world,) + self.shape)
