#!/usr/bin/env python
"""Round-2 parity probe (GPU box).  Writes gpurun_out/r2_parity_probe.json (+ lrf_svd_samples.npz).

1. Are torch's CUDA kernels on the reference's coarse path the algorithms we think they are?  Emulate, in numpy fp32,
   the summation ORDER of softmax(dim=2) (PersistentSoftmax.cuh softmax_warp_forward), softmax(dim=1) (sequential
   spatial kernel) and cumsum(dim=1) (ScanUtils.cuh: Sklansky scan on blocks of 2*num_threads_x with a carry into
   element 0) on torch's own CUDA elementwise results, and compare bit for bit.
2. Coarse solver vs the oracle (reference torch path on this GPU) on the same atten / u: selected pool index
   agreement, draws that sample a different correspondence, top-K overlap, and for each disagreement the oracle's
   rank of our winner.
3. LRF_batch: vote-tie rate at the PositionalEncoding scales, samples of (covariance, torch.svd V) for the sign study.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

from oracle import pose_oracle as PO  # noqa: E402
from unopose_b200 import model_utils as MU  # noqa: E402
from unopose_b200.synthetic import batch_clouds, matching_batch  # noqa: E402

dev = torch.device("cuda:0")
report = {}


def bits(x):
    return x.detach().cpu().contiguous().numpy().view(np.uint32)


# ----------------------------------------------------------------------------- 1. emulations
def emu_softmax_last(e_cpu):
    """e = exp(x - max) (n_rows, C) fp32 numpy -> row sums in softmax_warp_forward order."""
    n, C = e_cpu.shape
    p2 = 1
    while p2 < C:
        p2 *= 2
    W = min(p2, 32)
    iters = p2 // W
    pad = np.zeros((n, p2), np.float32)
    pad[:, :C] = e_cpu
    lane = np.zeros((n, W), np.float32)
    for it in range(iters):
        lane = (lane + pad[:, it * W:(it + 1) * W]).astype(np.float32)
    off = W // 2
    while off > 0:
        idx = np.arange(W) ^ off
        lane = (lane + lane[:, idx]).astype(np.float32)
        off //= 2
    return lane[:, 0]


def emu_seq_sum(e_cpu, axis):
    """sequential fp32 sum along `axis` (0 .. n-1)."""
    e = np.moveaxis(e_cpu, axis, 0)
    s = np.zeros(e.shape[1:], np.float32)
    for d in range(e.shape[0]):
        s = (s + e[d]).astype(np.float32)
    return s


def log_threads_x(num_rows, row_size):
    lx = 0
    while (1 << lx) < row_size:
        lx += 1
    ly = 0
    while (1 << ly) < num_rows:
        ly += 1
    lx = (9 + (lx - ly)) // 2
    return min(max(4, lx), 9)


def emu_cumsum(p_cpu):
    """ScanUtils.cuh tensor_kernel_scan_innermost_dim_impl on a (rows, L) fp32 array."""
    rows, L = p_cpu.shape
    lx = log_threads_x(rows, L)
    ntx = 1 << lx
    Wb = 2 * ntx
    out = np.empty_like(p_cpu)
    total = np.zeros(rows, np.float32)
    t = np.arange(ntx)
    for c0 in range(0, L, Wb):
        buf = np.zeros((rows, Wb), np.float32)
        n = min(Wb, L - c0)
        buf[:, :n] = p_cpu[:, c0:c0 + n]
        buf[:, 0] = (buf[:, 0] + total).astype(np.float32)
        for m in range(lx + 1):
            s = 1 << m
            a = ((t >> m) << (m + 1)) | s
            ti = a + (t % s)
            si = a - 1
            buf[:, ti] = (buf[:, ti] + buf[:, si]).astype(np.float32)
        out[:, c0:c0 + n] = buf[:, :n]
        total = buf[:, Wb - 1].copy()
    return out


def probe_emulations():
    r = {}
    d = matching_batch(3, 16, 196, 256)
    f1, f2 = torch.from_numpy(d["f1"]).to(dev), torch.from_numpy(d["f2"]).to(dev)
    x = PO.feature_similarity(f1, f2, "cosine", 0.1, True)              # (16,197,197)
    # softmax over the last dim
    sm2 = torch.softmax(x, dim=2)
    mx = x.max(dim=2, keepdim=True)[0]
    e = torch.exp(x - mx)
    s_emu = emu_softmax_last(e.cpu().numpy().reshape(-1, x.shape[2]))
    sm2_emu = (e.cpu().numpy().reshape(-1, x.shape[2]) / s_emu[:, None]).astype(np.float32)
    r["softmax_dim2_bit_mismatch"] = int((sm2_emu.view(np.uint32) != bits(sm2).reshape(sm2_emu.shape)).sum())
    # softmax over dim 1
    sm1 = torch.softmax(x, dim=1)
    mx1 = x.max(dim=1, keepdim=True)[0]
    e1 = torch.exp(x - mx1)
    s1_emu = emu_seq_sum(e1.cpu().numpy(), 1)                           # (16,197)
    sm1_emu = (e1.cpu().numpy() / s1_emu[:, None, :]).astype(np.float32)
    r["softmax_dim1_bit_mismatch_sequential"] = int((sm1_emu.view(np.uint32) != bits(sm1)).sum())
    r["softmax_elements"] = int(sm1.numel())
    # pow 1.5 vs x*sqrt(x)
    A = (sm2 * sm1)[:, 1:, 1:].contiguous()
    p_pow = A ** 1.5
    p_sq = A * torch.sqrt(A)
    r["pow15_vs_xsqrtx_bit_mismatch"] = int((bits(p_pow) != bits(p_sq)).sum())
    r["pow15_elements"] = int(A.numel())
    # cumsum for several batch sizes
    P = p_pow.reshape(16, -1)
    for rows in (1, 2, 3, 16):
        pr = P[:rows].contiguous()
        cs = torch.cumsum(pr, dim=1)
        emu = emu_cumsum(pr.cpu().numpy())
        r["cumsum_rows%d_bit_mismatch" % rows] = int((emu.view(np.uint32) != bits(cs)).sum())
        cs2 = torch.cumsum(pr, dim=1)
        r["cumsum_rows%d_run_to_run_mismatch" % rows] = int((bits(cs) != bits(cs2)).sum())
    big = torch.rand(200, 38416, device=dev) ** 8
    r["cumsum_rows200_bit_mismatch"] = int((emu_cumsum(big.cpu().numpy()).view(np.uint32) != bits(torch.cumsum(big, 1))).sum())
    # the division
    cs = torch.cumsum(P, dim=1)
    cdf = cs / (cs[:, -1].unsqueeze(1).contiguous() + 1e-8)
    csn = cs.cpu().numpy()
    den = (csn[:, -1] + np.float32(1e-8)).astype(np.float32)
    r["cdf_division_bit_mismatch"] = int(((csn / den[:, None]).astype(np.float32).view(np.uint32) != bits(cdf)).sum())
    return r


# ----------------------------------------------------------------------------- 2. coarse agreement
def coarse_case(seed, B, n, H, K, feat_dim=256):
    d = matching_batch(seed, B, n, feat_dim)
    T = {k: torch.from_numpy(v).to(dev) for k, v in d.items() if k in ("pts1", "pts2", "f1", "f2", "score", "R", "t")}
    atten = PO.feature_similarity(T["f1"], T["f2"], "cosine", 0.1, True)
    u = torch.rand(B, 3 * H, generator=torch.Generator().manual_seed(seed)).to(dev)
    R, t, s, m = MU._coarse(atten, T["score"], T["pts1"], T["pts2"], None, H, K, u=u, return_debug=True)
    rows = []
    # the oracle on the whole batch (torch's cumsum kernel depends on the number of rows) when it fits
    if B * H <= 400000:
        Ro, to, so, o = PO.coarse_pose(atten, T["score"], T["pts1"], T["pts2"], None, H, K, u=u, debug=True)
        per = [(Ro[b], to[b], so[b], {k: v[b] for k, v in o.items()}) for b in range(B)]
    else:
        per = []
        for b in range(B):
            Ro, to, so, o = PO.coarse_pose(atten[b:b + 1], T["score"][b:b + 1], T["pts1"][b:b + 1], T["pts2"][b:b + 1],
                                           None, H, K, u=u[b:b + 1], debug=True)
            per.append((Ro[0], to[0], so[0], {k: v[0] for k, v in o.items()}))
    for b in range(B):
        Ro, to, so, o = per[b]
        mi1, mi2 = m["idx1"][b].reshape(-1).long(), m["idx2"][b].reshape(-1).long()
        ddiff = (mi1 != o["idx1"]) | (mi2 != o["idx2"])
        hyp_diff = ddiff.reshape(H, 3).any(1)
        top_m, top_o = set(m["top"][b].tolist()), set(o["top"].tolist())
        mine, theirs = int(m["pool"][b]), int(o["pool"])
        row = dict(seed=seed, b=b, H=H, K=K, agree=mine == theirs, draws_differ=int(ddiff.sum()),
                   hyp_differ=int(hyp_diff.sum()), topk_common=len(top_m & top_o),
                   cdf_bits_differ=int((bits(m["cdf"][b]) != bits(o["cdf"])).sum()),
                   cdf_max_abs=float((m["cdf"][b] - o["cdf"]).abs().max()),
                   score_rel=float(abs(float(s[b]) - float(so)) / abs(float(so))),
                   rot_deg=float(PO.rotation_geodesic_deg(R[b], Ro)), t_rel=float(PO.relative_translation_error(t[b], to)))
        if mine != theirs:
            osc = o["scores"]
            order = torch.argsort(osc, descending=True)
            otop = o["top"][order].tolist()
            row["oracle_rank_of_our_winner"] = otop.index(mine) if mine in otop else -1
            if mine in otop:
                row["oracle_score_gap_rel"] = float((osc.max() - osc[order[otop.index(mine)]]) / osc.max())
            ms = m["scores"][b]
            mtop = m["top"][b].tolist()
            row["our_rank_of_oracle_winner"] = (torch.argsort(ms, descending=True).tolist().index(mtop.index(theirs))
                                                if theirs in mtop else -1)
            row["winner_hyp_differs_by_draw"] = bool(hyp_diff[mine]) or bool(hyp_diff[theirs])
            row["same_triplets"] = bool(torch.equal(o["idx1"].reshape(H, 3)[mine], o["idx1"].reshape(H, 3)[theirs]) and
                                        torch.equal(o["idx2"].reshape(H, 3)[mine], o["idx2"].reshape(H, 3)[theirs]))
        rows.append(row)
    return rows


def probe_coarse():
    rows = []
    for seed in range(100, 110):
        rows += coarse_case(seed, 4, 196, 5000, 300)
    for seed in range(200, 203):
        rows += coarse_case(seed, 16, 196, 5000, 300)
    rows += coarse_case(300, 8, 196, 1000, 300)
    rows += coarse_case(301, 8, 196, 6000, 300)
    rows += coarse_case(302, 8, 196, 20000, 300)
    rows += coarse_case(303, 4, 196, 50000, 300)
    rows += coarse_case(304, 2, 196, 100000, 300)
    rows += coarse_case(305, 1, 196, 100000, 300)
    summ = {}
    for r in rows:
        k = "H%d" % r["H"]
        s = summ.setdefault(k, dict(instances=0, agree=0, draws_differ=0, hyp_differ=0, cdf_bits_differ=0))
        s["instances"] += 1
        s["agree"] += int(r["agree"])
        for f in ("draws_differ", "hyp_differ", "cdf_bits_differ"):
            s[f] += r[f]
    return dict(summary=summ, disagreements=[r for r in rows if not r["agree"]], rows=rows)


# ----------------------------------------------------------------------------- 3. LRF sign ties
def probe_lrf():
    from unopose_b200.pointnet2 import pointnet2_utils as PU

    out = {}
    samples = {}
    for kind in ("surface", "ball"):
        pts = torch.from_numpy(batch_clouds(31, 8, 2048, kind)).to(dev).contiguous()
        for r, ns in ((0.1, 64), (0.2, 256)):
            idx, grouped = PU.ball_query_and_group(pts, pts, [(r, ns)])[0]          # grouped (B,3,N,ns)
            g = grouped.transpose(1, 2).contiguous()                                  # (B,N,3,ns)
            x = pts.unsqueeze(3) - g
            xxt = torch.einsum("bcnj,bcjm->bcnm", x, x.transpose(2, 3)) / ns
            _, sv, v = torch.svd(xxt)
            z = v[..., -1]
            h = (z.unsqueeze(2) @ x).squeeze(2)
            vote = (h > 1e-3).sum(-1) - (h < -1e-3).sum(-1)
            key = "%s_r%.1f_ns%d" % (kind, r, ns)
            out[key] = dict(centres=int(vote.numel()), tie=int((vote == 0).sum()), tie_rate=float((vote == 0).float().mean()),
                            near_tie_abs_le_1=float((vote.abs() <= 1).float().mean()),
                            rank_deficient=float((sv[..., 1] < 1e-9).float().mean()))
            tie = (vote == 0)
            sel = torch.nonzero(tie.reshape(-1)).flatten()[:20000]
            samples[key + "_cov"] = xxt.reshape(-1, 3, 3)[sel].cpu().numpy()
            samples[key + "_v"] = v.reshape(-1, 3, 3)[sel].cpu().numpy()
            samples[key + "_s"] = sv.reshape(-1, 3)[sel].cpu().numpy()
            allsel = torch.arange(0, vote.numel(), 7, device=dev)[:5000]
            samples[key + "_cov_all"] = xxt.reshape(-1, 3, 3)[allsel].cpu().numpy()
            samples[key + "_v_all"] = v.reshape(-1, 3, 3)[allsel].cpu().numpy()
    np.savez_compressed(os.path.join(OUT, "lrf_svd_samples.npz"), **samples)
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["emu", "coarse", "lrf"]
    if "emu" in which:
        report["emulation"] = probe_emulations()
        print(json.dumps(report["emulation"], indent=1))
    if "coarse" in which:
        report["coarse"] = probe_coarse()
        print(json.dumps(report["coarse"]["summary"], indent=1))
        print(json.dumps(report["coarse"]["disagreements"], indent=1))
    if "lrf" in which:
        report["lrf"] = probe_lrf()
        print(json.dumps(report["lrf"], indent=1))
    json.dump(report, open(os.path.join(OUT, "r2_parity_probe.json"), "w"), indent=1)
