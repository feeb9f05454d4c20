"""Multi-GPU partitioning of the hot path: one process per GPU, torch.distributed.

A. instance sharding   — instances are independent; each rank runs its slice, the only collective is an
                         all_gather of the (B_local, 13) result rows [R(9) | t(3) | score].
B. hypothesis sharding — one instance batch, H hypotheses split across ranks.  Every rank builds the same
                         CDF and takes its slice of the SAME uniform draws; fixed-size candidate lists (residual,
                         pool index, R, t) are all-gathered, the global top-K is re-selected by the same
                         kernel on every rank, each rank scores its slice of the kept list, the score table is
                         completed by an all_reduce(MAX) against -inf and the arg-max is taken.  No float is
                         ever combined across ranks, so the result is bit-identical to one GPU.

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
    so the SAME top-K kernel that a single GPU runs re-selects the global top-K (identical tie rule).
    (torch formulation of upk_unpack_candidates; used by the CPU/gloo tests of the host logic.)"""
    B = cand_resid.shape[0]
    dense = torch.full((B, n_hyp), float("inf"), dtype=cand_resid.dtype, device=cand_resid.device)
    dense.scatter_(1, cand_idx.long(), cand_resid)
    return dense


def compact_candidates(cand_resid, cand_idx):
    """The gathered candidate lists (world, B, kc) as the compact pool (B, world*kc) the merge selects from: residuals
    with padding records (idx < 0) moved behind everything, and the pool index of each entry.  Torch formulation of
    upk_unpack_candidates_compact (CPU/gloo tests of the host logic)."""
    W, B, kc = cand_resid.shape
    r = cand_resid.permute(1, 0, 2).reshape(B, W * kc).clone()
    i = cand_idx.permute(1, 0, 2).reshape(B, W * kc)
    r[i < 0] = float("inf")
    return r, i


def score_shard_range(K, rank, world):
    """Slice [k0, k1) of the kept list a rank scores: equal chunks of ceil(K / world), the tail ranks may be empty."""
    kmax = -(-K // world)
    k0 = min(K, rank * kmax)
    return k0, min(K, k0 + kmax)


def candidate_slots(H, K, world):
    """Records per instance in a rank's candidate list: min(K, largest slice)."""
    return min(K, max(shard_range(H, r, world)[1] - shard_range(H, r, world)[0] for r in range(world)))


class HypothesisShardedCoarse:
    """compute_coarse_Rt_overlap with the hypothesis pool split across the ranks of `group` (partitioning B).

    All buffers are allocated once (static addresses, so a solve can be captured into a CUDA graph together with its
    two collectives).  One solve = 8 kernel launches + 2 exchanges:
      assignment (replicated, bit-exact cluster kernel) -> my slice of the hypotheses -> top-K of the slice (in place,
      pitched) -> pack candidates -> EXCHANGE #1 (world x B x slots x 56 B) -> the gathered lists as a COMPACT pool of
      world x slots candidates per instance (compact order == pool-index order, so the same top-K kernel applies the
      same tie rule; no H-sized array is touched after the sampling) -> global top-K -> score my slice of the kept list
      -> EXCHANGE #2 over the (B, K) score table (every entry is written by exactly one rank; with NCCL an
      ALL_REDUCE(MAX) against -inf moves it unchanged, no float is ever combined across ranks) -> arg-max + gather.
    The result is identical on every rank and bit-identical to the single-GPU solver.

    exchange = "p2p": the two collectives become peer-memory exchanges fused into the kernels on either side of them
    (csrc/peer.cuh): pack_candidates stores this rank's list into every rank's gathered array over NVLink and publishes,
    unpack_candidates waits on its local flags; score_hypotheses stores its slice of the score table into every rank's
    table, select_best waits.  No NCCL launch, no +-inf fills of the exchanged arrays, 8 kernel launches per solve.
    "auto" picks p2p when the ranks can map each other's memory (peer.available) and every rank scores a non-empty
    slice of the kept list, else NCCL; "nccl" forces the collectives (gloo on CPU tensors)."""

    def __init__(self, B, N1, N2, H, K, device, group=None, with_score=True, exchange="auto"):
        from . import _lib as L

        self.L, self.lib = L, L.load()
        self.group = group
        self.rank, self.world = _world(group)
        self.B, self.N1, self.N2, self.H, self.K = B, N1, N2, int(H), int(K)
        self.dev = torch.device(device)
        self.h0, self.h1 = shard_range(self.H, self.rank, self.world)
        self.kl = min(self.K, self.h1 - self.h0)
        self.kc = candidate_slots(self.H, self.K, self.world)
        self.k0, self.k1 = score_shard_range(self.K, self.rank, self.world)
        f32 = lambda *shape: torch.empty(shape, dtype=torch.float32, device=self.dev)  # noqa: E731
        i32 = lambda *shape: torch.empty(shape, dtype=torch.int32, device=self.dev)  # noqa: E731
        self.ws = torch.empty(max(self.lib.upk_coarse_assignment_workspace_bytes(B, N1, N2), 256), dtype=torch.uint8,
                              device=self.dev)
        self.w1, self.w2, self.cdf = f32(B, N1), f32(B, N2), f32(B, N1 * N2)
        self.Rs, self.ts, self.resid = torch.zeros((B, self.H, 9), device=self.dev), torch.zeros((B, self.H, 3), device=self.dev), f32(B, self.H)
        self.loc = f32(B, max(self.h1 - self.h0, 1))
        self.top_l, self.top = i32(B, max(self.kl, 1)), i32(B, self.K)
        self.cand = f32(B, self.kc, 14)
        self.allc = f32(self.world, B, self.kc, 14)
        nc = self.world * self.kc                    # the compact candidate pool of the merge (world > 1)
        self.resid_c, self.Rs_c, self.ts_c, self.pool_c = f32(B, nc), f32(B, nc, 9), f32(B, nc, 3), i32(B, nc)
        self.scores = f32(B, self.K)
        self.R, self.t, self.sc, self.pool = f32(B, 3, 3), f32(B, 3), f32(B), i32(B)
        if exchange not in ("auto", "nccl", "p2p"):
            raise ValueError(exchange)
        self.px = None
        if self.world > 1 and exchange != "nccl":
            from . import peer as P

            every_rank_scores = all(score_shard_range(self.K, r, self.world)[1] > score_shard_range(self.K, r, self.world)[0]
                                    for r in range(self.world))
            ok = every_rank_scores and P.available(self.dev, group)
            if exchange == "p2p" and not ok:
                raise RuntimeError("exchange='p2p' needs peer-mappable CUDA devices under NCCL and K >= world")
            if ok:
                cand_bytes, score_bytes = self.allc.numel() * 4, self.scores.numel() * 4
                self.px = P.PeerExchange(2 * (cand_bytes + 256) + 2 * (score_bytes + 256) + 1024, self.dev, group)
                self.ch1 = self.px.reserve(cand_bytes)
                self.ch2 = self.px.reserve(score_bytes)
        self.exchange = "p2p" if self.px is not None else ("nccl" if self.world > 1 else "none")

    def run(self, atten, score, pts1, pts2, u):
        """atten (B,N1+1,N2+1), score (B,N1+N2)|None, pts (B,N,3), u (B,3H) identical on every rank (same seed).
        Inputs must be fp32 contiguous CUDA tensors.  Returns (R, t, score, pool_idx) — the solver's own buffers."""
        L, lib, B, N1, N2, H, K = self.L, self.lib, self.B, self.N1, self.N2, self.H, self.K
        s1 = s2 = None
        ld = 0
        if score is not None:
            if score.shape[1] - N2 != N2:
                raise RuntimeError("compute_coarse_Rt_overlap: score[:, N2:] must have N2 columns "
                                   "(the reference slices score[:, N2:], which requires N1 == N2)")
            ld = score.shape[1]
            s1, s2 = score, score[:, N2:]
        st = L.stream_ptr(pts1)
        with torch.cuda.device(self.dev):
            if self.world == 1:
                L.check(lib.upk_fill_f32(L.ptr(self.resid), self.resid.numel(), float("inf"), st), "fill")
            L.check(lib.upk_coarse_assignment(L.ptr(atten), L.ptr(s1), ld, s2.data_ptr() if s2 is not None else None, ld,
                                              B, N1, N2, L.ptr(self.ws), self.ws.numel(), L.ptr(self.w1), L.ptr(self.w2),
                                              L.ptr(self.cdf), st), "coarse_assignment")
            L.check(lib.upk_sample_hypotheses(L.ptr(self.cdf), L.ptr(u), L.ptr(pts1), L.ptr(pts2), B, N1, N2, H, self.h0,
                                              self.h1, None, None, L.ptr(self.Rs), L.ptr(self.ts), L.ptr(self.resid), st),
                    "sample_hypotheses")
            if self.world == 1:
                L.check(lib.upk_topk_smallest(L.ptr(self.resid), B, H, K, L.ptr(self.top), st), "topk_global")
                L.check(lib.upk_score_hypotheses(L.ptr(pts1), L.ptr(pts2), L.ptr(self.w1), L.ptr(self.Rs), L.ptr(self.ts),
                                                 L.ptr(self.top), B, N1, N2, H, K, 0, K, L.ptr(self.scores), st),
                        "score_hypotheses")
                L.check(lib.upk_select_best(L.ptr(self.scores), L.ptr(self.top), L.ptr(self.Rs), L.ptr(self.ts), B, H, K,
                                            L.ptr(self.R), L.ptr(self.t), L.ptr(self.sc), L.ptr(self.pool), st), "select_best")
                return self.R, self.t, self.sc, self.pool
            # ---- my slice's K best (the top-K kernel on the slice of the pool array, row pitch H)
            if self.kl > 0:
                L.check(lib.upk_topk_smallest_ld(self.resid.data_ptr() + 4 * self.h0, B, self.h1 - self.h0, H, self.kl,
                                                 L.ptr(self.top_l), st), "topk_local")
            nc = self.world * self.kc
            if self.px is not None:                                                            # exchange #1, peer memory
                ch, off, slab = self.ch1
                L.check(lib.upk_pack_candidates_peer(L.ptr(self.resid), L.ptr(self.Rs), L.ptr(self.ts), L.ptr(self.top_l),
                                                     B, H, self.h0, self.kl, self.kc, self.px.ref, off, slab, ch, st),
                        "pack_candidates_peer")
                L.check(lib.upk_unpack_candidates_compact_peer(self.px.ref, off, slab, ch, B, self.kc, L.ptr(self.resid_c),
                                                               L.ptr(self.Rs_c), L.ptr(self.ts_c), L.ptr(self.pool_c), st),
                        "unpack_candidates_compact_peer")
            else:
                L.check(lib.upk_pack_candidates(L.ptr(self.resid), L.ptr(self.Rs), L.ptr(self.ts), L.ptr(self.top_l), B, H,
                                                self.h0, self.kl, self.kc, L.ptr(self.cand), st), "pack_candidates")
                dist.all_gather_into_tensor(self.allc, self.cand, group=self.group)           # collective #1
                L.check(lib.upk_unpack_candidates_compact(L.ptr(self.allc), self.world, B, self.kc, L.ptr(self.resid_c),
                                                          L.ptr(self.Rs_c), L.ptr(self.ts_c), L.ptr(self.pool_c), st),
                        "unpack_candidates_compact")
            # ---- global top-K over the world * kc candidates (compact order == pool-index order: same tie rule)
            L.check(lib.upk_topk_smallest(L.ptr(self.resid_c), B, nc, K, L.ptr(self.top), st), "topk_global")
            if self.px is not None:                                                            # exchange #2, peer memory
                ch, off, slab = self.ch2
                L.check(lib.upk_score_hypotheses_peer(L.ptr(pts1), L.ptr(pts2), L.ptr(self.w1), L.ptr(self.Rs_c),
                                                      L.ptr(self.ts_c), L.ptr(self.top), B, N1, N2, nc, K, self.k0, self.k1,
                                                      self.px.ref, off, slab, ch, st), "score_hypotheses_peer")
                L.check(lib.upk_select_best_peer(self.px.ref, off, slab, ch, L.ptr(self.top), L.ptr(self.Rs_c),
                                                 L.ptr(self.ts_c), L.ptr(self.pool_c), B, nc, K, L.ptr(self.R), L.ptr(self.t),
                                                 L.ptr(self.sc), L.ptr(self.pool), st), "select_best_peer")
                return self.R, self.t, self.sc, self.pool
            L.check(lib.upk_fill_f32(L.ptr(self.scores), self.scores.numel(), float("-inf"), st), "fill")
            L.check(lib.upk_score_hypotheses(L.ptr(pts1), L.ptr(pts2), L.ptr(self.w1), L.ptr(self.Rs_c), L.ptr(self.ts_c),
                                             L.ptr(self.top), B, N1, N2, nc, K, self.k0, self.k1, L.ptr(self.scores), st),
                    "score_hypotheses")
            dist.all_reduce(self.scores, op=dist.ReduceOp.MAX, group=self.group)                 # collective #2
            L.check(lib.upk_select_best_map(L.ptr(self.scores), L.ptr(self.top), L.ptr(self.Rs_c), L.ptr(self.ts_c),
                                            L.ptr(self.pool_c), B, nc, K, L.ptr(self.R), L.ptr(self.t), L.ptr(self.sc),
                                            L.ptr(self.pool), st), "select_best_map")
        return self.R, self.t, self.sc, self.pool


def coarse_pose_hypothesis_sharded(atten, score, pts1, pts2, n_proposal1, n_proposal2, u, group=None, exchange="nccl"):
    """One-shot form of HypothesisShardedCoarse (allocates its buffers per call).  `u` (B, 3*n_proposal1) must be
    identical on every rank (draw it from the same seed).  Returns (R, t, score, pool_idx), identical on every rank and
    bit-identical to the single-GPU solver."""
    B, N1, _ = pts1.shape
    atten, pts1, pts2, u = (x.float().contiguous() for x in (atten, pts1, pts2, u))
    if score is not None:
        score = score.float().contiguous()
    solver = HypothesisShardedCoarse(B, N1, pts2.shape[1], n_proposal1, n_proposal2, pts1.device, group, exchange=exchange)
    R, t, sc, pool = solver.run(atten, score, pts1, pts2, u)
    return R.clone(), t.clone(), sc.clone(), pool.clone()
