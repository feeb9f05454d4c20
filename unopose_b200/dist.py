"""Multi-GPU partitioning of the hot path: one process per GPU, torch.distributed.

A. instance sharding   — instances are independent; each rank runs its slice, the only collective is an
                         all_gather of the (B_local, 13) result rows [R(9) | t(3) | score].
B. hypothesis sharding — one instance batch, H hypotheses split across ranks.  Every rank builds the same
                         CDF and takes its slice of the SAME uniform draws; candidates (residual, pool
                         index, R, t) are all-gathered, the global top-K is re-selected by the same
                         kernel on every rank, each rank scores its slice of the kept list, the scores are
                         all-gathered and the arg-max is taken.  No float is ever reduced across ranks, so
                         the result is bit-identical to one GPU.

The collectives work on whatever device the tensors live on (NCCL for CUDA, gloo for the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced slice [begin, end) of range(n) for `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def _world(group=None):
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def all_gather_cat(x, dim=0, group=None):
    """all_gather of equally-shaped tensors, concatenated along `dim` (identity when not distributed)."""
    rank, world = _world(group)
    if world == 1:
        return x
    parts = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(parts, x.contiguous(), group=group)
    return torch.cat(parts, dim=dim)


def all_gather_ragged(x, sizes, dim=0, group=None):
    """all_gather of tensors whose size along `dim` is sizes[rank] (padded to the max, then trimmed)."""
    rank, world = _world(group)
    if world == 1:
        return x
    mx = max(sizes)
    pad_shape = list(x.shape)
    pad_shape[dim] = mx
    buf = x.new_zeros(pad_shape)
    buf.narrow(dim, 0, x.shape[dim]).copy_(x)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p.narrow(dim, 0, s) for p, s in zip(parts, sizes)], dim=dim)


def pack_results(R, t, score):
    """(B,3,3),(B,3),(B,) -> (B,13) rows, the unit exchanged between ranks."""
    return torch.cat([R.reshape(R.shape[0], 9), t, score.unsqueeze(1)], dim=1)


def unpack_results(rows):
    return rows[:, :9].reshape(-1, 3, 3), rows[:, 9:12], rows[:, 12]


def gather_results(R, t, score, counts=None, group=None):
    """Instance sharding: every rank contributes its instances' results; returns the full batch in rank
    order.  `counts[r]` = number of instances on rank r (needed only when they differ)."""
    rows = pack_results(R, t, score)
    rank, world = _world(group)
    if world == 1:
        return unpack_results(rows)
    if counts is None:
        return unpack_results(all_gather_cat(rows, 0, group))
    return unpack_results(all_gather_ragged(rows, list(counts), 0, group))


def merge_topk_candidates(cand_resid, cand_idx, n_hyp):
    """Scatter gathered candidates (B, M) back into a dense (B, n_hyp) residual array filled with +inf,
    so the SAME top-K kernel that a single GPU runs re-selects the global top-K (identical tie rule)."""
    B = cand_resid.shape[0]
    dense = torch.full((B, n_hyp), float("inf"), dtype=cand_resid.dtype, device=cand_resid.device)
    dense.scatter_(1, cand_idx.long(), cand_resid)
    return dense


def coarse_pose_hypothesis_sharded(atten, score, pts1, pts2, n_proposal1, n_proposal2, u, group=None):
    """compute_coarse_Rt_overlap with the hypotheses split across the ranks of `group` (partitioning B).
    `u` (B, 3*n_proposal1) must be identical on every rank (draw it from the same seed).  Returns
    (R, t, score, pool_idx), identical on every rank and bit-identical to the single-GPU solver."""
    import ctypes  # noqa: F401

    from . import _lib as L

    rank, world = _world(group)
    B, N1, _ = pts1.shape
    N2 = pts2.shape[1]
    H, K = int(n_proposal1), int(n_proposal2)
    dev = pts1.device
    lib = L.load()
    atten, pts1, pts2, u = (x.float().contiguous() for x in (atten, pts1, pts2, u))
    s1 = s2 = None
    ld = 0
    if score is not None:
        score = score.float().contiguous()
        ld = score.shape[1]
        s1, s2 = score, score[:, N2:]
    st = lambda: L.stream_ptr(pts1)
    f32 = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        # replicated: masks + CDF
        ws = torch.empty(max(lib.upk_coarse_assignment_workspace_bytes(B, N1, N2), 256), dtype=torch.uint8, device=dev)
        w1, w2, cdf = f32(B, N1), f32(B, N2), f32(B, N1 * N2)
        L.check(lib.upk_coarse_assignment(L.ptr(atten), L.ptr(s1), ld, s2.data_ptr() if s2 is not None else None, ld,
                                          B, N1, N2, L.ptr(ws), ws.numel(), L.ptr(w1), L.ptr(w2), L.ptr(cdf), st()),
                "coarse_assignment")
        # my slice of the hypothesis pool
        h0, h1 = shard_range(H, rank, world)
        Rs, ts = torch.zeros((B, H, 9), device=dev), torch.zeros((B, H, 3), device=dev)
        resid = torch.full((B, H), float("inf"), device=dev)
        L.check(lib.upk_sample_hypotheses(L.ptr(cdf), L.ptr(u), L.ptr(pts1), L.ptr(pts2), B, N1, N2, H, h0, h1,
                                          None, None, L.ptr(Rs), L.ptr(ts), L.ptr(resid), st()), "sample_hypotheses")
        if world > 1:
            # local candidates: the K smallest of my slice (the global top-K is inside the union)
            sizes = [min(K, shard_range(H, r, world)[1] - shard_range(H, r, world)[0]) for r in range(world)]
            kl = sizes[rank]
            loc = resid[:, h0:h1].contiguous()
            top_l = torch.empty((B, kl), dtype=torch.int32, device=dev)
            L.check(lib.upk_topk_smallest(L.ptr(loc), B, h1 - h0, kl, L.ptr(top_l), st()), "topk_local")
            idx = top_l.long() + h0
            cand = torch.cat([torch.gather(resid, 1, idx).unsqueeze(2), idx.to(torch.float32).unsqueeze(2),
                              torch.gather(Rs, 1, idx.unsqueeze(2).expand(-1, -1, 9)),
                              torch.gather(ts, 1, idx.unsqueeze(2).expand(-1, -1, 3))], dim=2)  # (B,kl,14)
            allc = all_gather_ragged(cand, sizes, dim=1, group=group)                          # collective #1
            gi = allc[:, :, 1].long()
            resid = merge_topk_candidates(allc[:, :, 0].contiguous(), gi, H)
            Rs.scatter_(1, gi.unsqueeze(2).expand(-1, -1, 9), allc[:, :, 2:11].contiguous())
            ts.scatter_(1, gi.unsqueeze(2).expand(-1, -1, 3), allc[:, :, 11:14].contiguous())
        top = torch.empty((B, K), dtype=torch.int32, device=dev)
        L.check(lib.upk_topk_smallest(L.ptr(resid), B, H, K, L.ptr(top), st()), "topk_global")
        k0, k1 = shard_range(K, rank, world)
        scores = torch.zeros((B, K), device=dev)
        L.check(lib.upk_score_hypotheses(L.ptr(pts1), L.ptr(pts2), L.ptr(w1), L.ptr(Rs), L.ptr(ts), L.ptr(top), B,
                                         N1, N2, H, K, k0, k1, L.ptr(scores), st()), "score_hypotheses")
        if world > 1:
            ksz = [shard_range(K, r, world)[1] - shard_range(K, r, world)[0] for r in range(world)]
            scores = all_gather_ragged(scores[:, k0:k1].contiguous(), ksz, dim=1, group=group)   # collective #2
        R, t, sc = f32(B, 3, 3), f32(B, 3), f32(B)
        pool = torch.empty((B,), dtype=torch.int32, device=dev)
        L.check(lib.upk_select_best(L.ptr(scores.contiguous()), L.ptr(top), L.ptr(Rs), L.ptr(ts), B, H, K, L.ptr(R),
                                    L.ptr(t), L.ptr(sc), L.ptr(pool), st()), "select_best")
    return R, t, sc, pool
