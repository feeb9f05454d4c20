"""torchrun worker: hypothesis sharding over NCCL must reproduce the single-GPU result bit for bit."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pose_oracle as PO  # noqa: E402
from unopose_b200 import model_utils as MU  # noqa: E402
from unopose_b200.dist import coarse_pose_hypothesis_sharded, gather_results, shard_range  # noqa: E402
from util_clouds import matching_batch  # noqa: E402

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
ok = True
for H, K in ((1000, 300), (5000, 300), (20000, 300)):
    B, n = 4, 196
    d = {k: torch.from_numpy(v).to(dev) for k, v in matching_batch(7, B, n, 128).items() if k in ("pts1", "pts2", "f1", "f2", "score")}
    atten = PO.feature_similarity(d["f1"], d["f2"], "cosine", 0.1, True)
    u = torch.rand(B, 3 * H, generator=torch.Generator().manual_seed(H)).to(dev)   # identical on every rank
    R, t, s, m = MU._coarse(atten, d["score"], d["pts1"], d["pts2"], None, H, K, u=u, return_debug=True)
    R2, t2, s2, pool2 = coarse_pose_hypothesis_sharded(atten, d["score"], d["pts1"], d["pts2"], H, K, u)
    ok = ok and torch.equal(R, R2) and torch.equal(t, t2) and torch.equal(s, s2) and torch.equal(m["pool"], pool2)
    # instance sharding: each rank solves its slice, results gathered
    b0, b1 = shard_range(B, rank, world)
    Rl, tl, sl = MU._coarse(atten[b0:b1], d["score"][b0:b1], d["pts1"][b0:b1], d["pts2"][b0:b1], None, H, K, u=u[b0:b1])
    counts = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    Rg, tg, sg = gather_results(Rl, tl, sl, counts=counts)
    ok = ok and torch.equal(Rg, R) and torch.equal(tg, t) and torch.equal(sg, s)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARD_OK" if int(flag.item()) == 1 else "SHARD_MISMATCH")
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
