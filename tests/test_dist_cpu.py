"""CPU: the N>1 host logic (unopose_b200/dist.py) under world_size-2 gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unopose_b200.dist import (all_gather_ragged, gather_results, merge_topk_candidates, pack_results, shard_range,
                               unpack_results)


def test_shard_range_partitions():
    for n in (0, 1, 5, 16, 300, 5000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    R, t, s = torch.randn(5, 3, 3), torch.randn(5, 3), torch.randn(5)
    r2, t2, s2 = unpack_results(pack_results(R, t, s))
    assert torch.equal(R, r2) and torch.equal(t, t2) and torch.equal(s, s2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        # ---- A. instance sharding: 7 instances over 2 ranks (ragged 4 + 3)
        B = 7
        R, t, s = torch.randn(B, 3, 3, generator=g), torch.randn(B, 3, generator=g), torch.randn(B, generator=g)
        b0, b1 = shard_range(B, rank, world)
        counts = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
        Rg, tg, sg = gather_results(R[b0:b1], t[b0:b1], s[b0:b1], counts=counts)
        okA = torch.equal(Rg, R) and torch.equal(tg, t) and torch.equal(sg, s)
        # equal split (8 instances) without counts
        R8 = torch.randn(8, 3, 3, generator=g)
        Rg8, _, _ = gather_results(R8[rank * 4:(rank + 1) * 4], torch.zeros(4, 3), torch.zeros(4))
        okA = okA and torch.equal(Rg8, R8)
        # ---- B. hypothesis sharding: the union of per-shard top-K contains the global top-K; re-selecting
        #         on the scattered dense array reproduces the single-process selection exactly
        Bn, H, K = 3, 1000, 37
        resid = torch.rand(Bn, H, generator=g)
        resid[:, 100:110] = resid[:, 5:6]  # ties
        h0, h1 = shard_range(H, rank, world)
        sizes = [min(K, shard_range(H, r, world)[1] - shard_range(H, r, world)[0]) for r in range(world)]
        loc_idx = torch.topk(resid[:, h0:h1], sizes[rank], dim=1, largest=False)[1] + h0
        cand = torch.stack([torch.gather(resid, 1, loc_idx), loc_idx.float()], dim=2)
        allc = all_gather_ragged(cand, sizes, dim=1)
        dense = merge_topk_candidates(allc[:, :, 0].contiguous(), allc[:, :, 1].long(), H)
        # the K smallest values of the dense (scattered) array equal those of the full array
        mine = torch.sort(torch.topk(dense, K, dim=1, largest=False)[0], dim=1)[0]
        full = torch.sort(torch.topk(resid, K, dim=1, largest=False)[0], dim=1)[0]
        okB = torch.equal(mine, full) and bool((dense[dense != float("inf")] >= 0).all())
        out[rank] = bool(okA and okB)
    finally:
        dist.destroy_process_group()


def test_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_score_and_candidate_partitions():
    from unopose_b200.dist import candidate_slots, score_shard_range

    for K in (1, 30, 300, 301):
        for world in (1, 2, 3, 4, 8):
            spans = [score_shard_range(K, r, world) for r in range(world)]
            assert spans[0][0] == 0 and max(e for _, e in spans) == K
            covered = sorted(k for b, e in spans for k in range(b, e))
            assert covered == list(range(K))                       # every kept hypothesis scored by exactly one rank
            assert len({e - b for b, e in spans if e > b} | {0}) <= 3
    assert candidate_slots(5000, 300, 8) == 300 and candidate_slots(1000, 300, 8) == 125 and candidate_slots(100, 300, 1) == 100


def _topk_smallest_lower_index_first(vals, K):
    """The tie rule of upk_topk_smallest: the K smallest, ties at the K-th value to the lower index, ascending index."""
    out = []
    for row in vals:
        order = sorted(range(len(row)), key=lambda i: (float(row[i]), i))[:K]
        out.append(sorted(order))
    return torch.tensor(out)


def test_compact_merge_selects_what_the_dense_merge_selects():
    """Hypothesis sharding, round 2: the global top-K runs on the compact pool of world*kc candidates instead of the dense
    H-sized array.  Compact order == pool-index order (every rank's list ascending, ranks own ascending slices), so
    the same top-K rule selects the same hypotheses in the same order — including exact ties and padded lists."""
    from unopose_b200.dist import candidate_slots, compact_candidates

    g = torch.Generator().manual_seed(5)
    for world, H, K in ((2, 1000, 37), (3, 500, 300), (8, 2000, 300), (4, 64, 30)):
        B = 2
        resid = torch.rand(B, H, generator=g)
        resid[:, 40:60] = resid[:, 3:4]                      # a block of exact ties straddling nothing in particular
        resid[:, H // 2 - 2:H // 2 + 2] = resid[:, 3:4]      # ... and across a slice boundary (world = 2)
        kc = candidate_slots(H, K, world)
        cr = torch.full((world, B, kc), float("inf"))
        ci = torch.full((world, B, kc), -1, dtype=torch.long)
        for r in range(world):
            h0, h1 = shard_range(H, r, world)
            kl = min(K, h1 - h0)
            loc = _topk_smallest_lower_index_first(resid[:, h0:h1], kl) + h0          # ascending pool index
            cr[r, :, :kl] = torch.gather(resid, 1, loc)
            ci[r, :, :kl] = loc
        rc, ic = compact_candidates(cr, ci)
        top_c = _topk_smallest_lower_index_first(rc, K)                              # indices into the compact pool
        sel_compact = torch.gather(ic, 1, top_c)
        sel_dense = _topk_smallest_lower_index_first(resid, K)                       # what one GPU selects
        assert torch.equal(sel_compact, sel_dense), (world, H, K)
        assert bool((sel_compact >= 0).all())
