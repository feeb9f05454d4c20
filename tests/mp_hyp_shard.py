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
from unopose_b200 import peer as P  # noqa: E402
from unopose_b200.dist import (HypothesisShardedCoarse, coarse_pose_hypothesis_sharded, gather_results, pack_results,  # noqa: E402
                               shard_range)
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

# ---- the same exchanges through peer memory (csrc/peer.cuh): repeated solves on ONE solver (epoch parity, slab reuse),
#      eager and as a replayed CUDA graph, must reproduce the single-GPU result bit for bit every time
p2p = P.available(dev)
if p2p:
    B, n, H, K = 4, 196, 5000, 300
    d = {k: torch.from_numpy(v).to(dev) for k, v in matching_batch(9, B, n, 128).items() if k in ("pts1", "pts2", "f1", "f2", "score")}
    atten = PO.feature_similarity(d["f1"], d["f2"], "cosine", 0.1, True).contiguous()
    solver = HypothesisShardedCoarse(B, n, n, H, K, dev, exchange="p2p")
    ok = ok and solver.exchange == "p2p"
    u = torch.empty(B, 3 * H, device=dev)
    gen = torch.Generator().manual_seed(123)

    def check(tag):
        global ok
        R, t, s, m = MU._coarse(atten, d["score"], d["pts1"], d["pts2"], None, H, K, u=u, return_debug=True)
        same = (torch.equal(R, solver.R) and torch.equal(t, solver.t) and torch.equal(s, solver.sc)
                and torch.equal(m["pool"], solver.pool))
        if not same:
            print("rank", rank, "p2p mismatch at", tag, flush=True)
        ok = ok and same

    for it in range(5):
        u.copy_(torch.rand(B, 3 * H, generator=gen))
        solver.run(atten, d["score"], d["pts1"], d["pts2"], u)
        check("eager %d" % it)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        solver.run(atten, d["score"], d["pts1"], d["pts2"], u)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        solver.run(atten, d["score"], d["pts1"], d["pts2"], u)
    for it in range(4):
        u.copy_(torch.rand(B, 3 * H, generator=gen))
        g.replay()
        check("graph %d" % it)
    # a rank that runs ahead must not overwrite a slab a slower peer still reads: skew the ranks
    for it in range(6):
        u.copy_(torch.rand(B, 3 * H, generator=gen))
        if it % 2 == rank % 2:
            torch.cuda._sleep(20_000_000)     # ~10 ms of device time on this rank's stream
        g.replay()
        check("skewed %d" % it)
    ok = ok and not solver.px.timed_out()
    # result rows of instance sharding through the generic peer all-gather
    ag = P.PeerAllGather((3, 13), torch.float32, dev)
    for it in range(3):
        rows = torch.full((3, 13), float(10 * it + rank), device=dev) + torch.arange(13, device=dev)
        got = ag.gather(rows)
        exp = torch.stack([torch.full((3, 13), float(10 * it + r), device=dev) + torch.arange(13, device=dev) for r in range(world)])
        ok = ok and torch.equal(got, exp)
    ok = ok and not ag.px.timed_out()
    if rank == 0:
        print("P2P_CHECKED", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARD_OK" if int(flag.item()) == 1 else "SHARD_MISMATCH")
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
