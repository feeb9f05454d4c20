"""GPU: hypothesis-sharded coarse solver (unopose_b200/dist.py) is bit-identical to the single-GPU solver.
world=1 always; world=2 under torchrun/NCCL when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest
import torch

from oracle import pose_oracle as PO
from util_clouds import matching_batch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_equals_fused_world1(cuda):
    from unopose_b200 import model_utils as MU
    from unopose_b200.dist import coarse_pose_hypothesis_sharded

    B, n, H, K = 3, 196, 5000, 300
    d = {k: torch.from_numpy(v).to(cuda) for k, v in matching_batch(5, B, n, 128).items() if k in ("pts1", "pts2", "f1", "f2", "score")}
    atten = PO.feature_similarity(d["f1"], d["f2"], "cosine", 0.1, True)
    u = torch.rand(B, 3 * H, device=cuda)
    R, t, s, m = MU._coarse(atten, d["score"], d["pts1"], d["pts2"], None, H, K, u=u, return_debug=True)
    R2, t2, s2, pool2 = coarse_pose_hypothesis_sharded(atten, d["score"], d["pts1"], d["pts2"], H, K, u)
    assert torch.equal(R, R2) and torch.equal(t, t2) and torch.equal(s, s2) and torch.equal(m["pool"], pool2)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_world2_nccl():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "mp_hyp_shard.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARD_OK" in r.stdout
    # the peer-memory exchange is exercised wherever the two GPUs can map each other's memory (NVLink / NVSwitch boxes)
    print("peer-memory exchange checked:", "P2P_CHECKED" in r.stdout)


@pytest.mark.parametrize("world,H", [(2, 5000), (3, 1000), (8, 20000)])
def test_compact_merge_emulated_ranks_on_one_gpu(cuda, world, H):
    """The stage-wise kernels of hypothesis sharding (pitched local top-K, pack, compact unpack, global top-K on the
    compact pool, sliced scoring, mapped selection) run for `world` emulated ranks one after the other on ONE GPU must
    reproduce the single-GPU solver bit for bit — the same data flow as dist.HypothesisShardedCoarse without a
    process group (so it is exercised on single-GPU boxes too)."""
    from unopose_b200 import _lib as L
    from unopose_b200 import model_utils as MU
    from unopose_b200.dist import candidate_slots, score_shard_range, shard_range

    lib = L.load()
    B, n, K = 3, 196, 300
    d = {k: torch.from_numpy(v).to(cuda) for k, v in matching_batch(21 + world, B, n, 128).items()
         if k in ("pts1", "pts2", "f1", "f2", "score")}
    atten = PO.feature_similarity(d["f1"], d["f2"], "cosine", 0.1, True).contiguous()
    u = torch.rand(B, 3 * H, generator=torch.Generator().manual_seed(H)).to(cuda)
    R0, t0, s0, m0 = MU._coarse(atten, d["score"], d["pts1"], d["pts2"], None, H, K, u=u, return_debug=True)
    st = torch.cuda.current_stream().cuda_stream
    f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=cuda)  # noqa: E731
    i32 = lambda *s: torch.empty(s, dtype=torch.int32, device=cuda)  # noqa: E731
    ws = torch.empty(max(lib.upk_coarse_assignment_workspace_bytes(B, n, n), 256), dtype=torch.uint8, device=cuda)
    w1, w2, cdf = f32(B, n), f32(B, n), f32(B, n * n)
    score = d["score"].contiguous()
    L.check(lib.upk_coarse_assignment(atten.data_ptr(), score.data_ptr(), 2 * n, score[:, n:].data_ptr(), 2 * n, B, n, n,
                                      ws.data_ptr(), ws.numel(), w1.data_ptr(), w2.data_ptr(), cdf.data_ptr(), st), "assign")
    kc = candidate_slots(H, K, world)
    allc = f32(world, B, kc, 14)
    for r in range(world):                                        # each "rank": its slice, its K best, its packed list
        h0, h1 = shard_range(H, r, world)
        kl = min(K, h1 - h0)
        Rs, ts, resid = torch.zeros(B, H, 9, device=cuda), torch.zeros(B, H, 3, device=cuda), f32(B, H)
        L.check(lib.upk_sample_hypotheses(cdf.data_ptr(), u.data_ptr(), d["pts1"].data_ptr(), d["pts2"].data_ptr(), B, n, n, H,
                                          h0, h1, None, None, Rs.data_ptr(), ts.data_ptr(), resid.data_ptr(), st), "sample")
        top_l = i32(B, kl)
        L.check(lib.upk_topk_smallest_ld(resid.data_ptr() + 4 * h0, B, h1 - h0, H, kl, top_l.data_ptr(), st), "topk_ld")
        ref_l = torch.topk(resid[:, h0:h1], kl, dim=1, largest=False, sorted=True)[0]
        assert torch.equal(torch.sort(torch.gather(resid[:, h0:h1], 1, top_l.long()), 1)[0], ref_l)
        L.check(lib.upk_pack_candidates(resid.data_ptr(), Rs.data_ptr(), ts.data_ptr(), top_l.data_ptr(), B, H, h0, kl, kc,
                                        allc[r].data_ptr(), st), "pack")
    nc = world * kc
    resid_c, Rs_c, ts_c, pool_c = f32(B, nc), f32(B, nc, 9), f32(B, nc, 3), i32(B, nc)
    L.check(lib.upk_unpack_candidates_compact(allc.data_ptr(), world, B, kc, resid_c.data_ptr(), Rs_c.data_ptr(),
                                              ts_c.data_ptr(), pool_c.data_ptr(), st), "unpack_compact")
    top = i32(B, K)
    L.check(lib.upk_topk_smallest(resid_c.data_ptr(), B, nc, K, top.data_ptr(), st), "topk")
    assert torch.equal(torch.gather(pool_c, 1, top.long()), m0["top"].to(torch.int32))        # same kept list, same order
    scores = torch.full((B, K), float("-inf"), device=cuda)
    for r in range(world):                                        # each "rank" scores its slice of the kept list
        k0, k1 = score_shard_range(K, r, world)
        if k1 > k0:
            L.check(lib.upk_score_hypotheses(d["pts1"].data_ptr(), d["pts2"].data_ptr(), w1.data_ptr(), Rs_c.data_ptr(),
                                             ts_c.data_ptr(), top.data_ptr(), B, n, n, nc, K, k0, k1, scores.data_ptr(), st),
                    "score")
    R, t, sc, pool = f32(B, 3, 3), f32(B, 3), f32(B), i32(B)
    L.check(lib.upk_select_best_map(scores.data_ptr(), top.data_ptr(), Rs_c.data_ptr(), ts_c.data_ptr(), pool_c.data_ptr(), B,
                                    nc, K, R.data_ptr(), t.data_ptr(), sc.data_ptr(), pool.data_ptr(), st), "select")
    torch.cuda.synchronize()
    assert torch.equal(R, R0) and torch.equal(t, t0) and torch.equal(sc, s0) and torch.equal(pool, m0["pool"].to(torch.int32))
