"""Dev probe: does torch's symmetric memory rendezvous work on this box (peer pointers for NVLink stores)?
torchrun --nproc-per-node 2 scripts/dev/symm_probe.py"""
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
print(rank, "can_access_peer", [torch.cuda.can_device_access_peer(local, j) for j in range(world) if j != local], flush=True)
try:
    import torch.distributed._symmetric_memory as symm

    t = symm.empty(1024, dtype=torch.float32, device=torch.device("cuda", local))
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "rendezvous ok; buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in hdl.signal_pad_ptrs],
          "multicast_ptr", getattr(hdl, "multicast_ptr", None), flush=True)
    t.fill_(float(rank + 1))
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
    print(rank, "peer read", float(peer[0]), flush=True)
    peer[1:3] = 100.0 + rank       # store into the peer's buffer over NVLink
    hdl.barrier()
    print(rank, "after peer store, local[0:4] =", t[:4].tolist(), flush=True)
except Exception as ex:  # noqa: BLE001
    import traceback

    traceback.print_exc()
    print(rank, "symmetric memory FAILED:", repr(ex)[:300], flush=True)
dist.barrier()
dist.destroy_process_group()
