"""A/B of the TMA-fed fine passes (UPK_FINE_TMA=1, default) against the register-streaming kernels (=0):
bit-identity of every intermediate (w1, w2, soft, asum) and of R / t / score, and the stage time as a CUDA graph.

    python scripts/dev/fine_tma_ab.py            # parent: runs both children and compares
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def child(out_path):
    import torch
    from unopose_b200 import model_utils as MU
    from unopose_b200.pipeline import HotPathConfig, synthetic_inputs

    cfg = HotPathConfig()
    dev = torch.device("cuda:0")
    B = int(os.environ.get("AB_BATCH", "16"))
    inp = synthetic_inputs(7, B, cfg, device=dev)
    with torch.no_grad():
        atten, stats = MU.compute_feature_similarity(inp["f_f1"], inp["f_f2"], "cosine", cfg.temp, True, return_stats=True)
        R, t, sc, dbg = MU._fine(atten, inp["f_score"], inp["f_pts1"], inp["f_pts2"], None, cfg.dis_thres, 0.001,
                                 return_debug=True, stats=stats)
        # also the path without the fused statistics (k_fine_stats + labels + rows)
        R2, t2, sc2, dbg2 = MU._fine(atten, inp["f_score"], inp["f_pts1"], inp["f_pts2"], None, cfg.dis_thres, 0.001,
                                     return_debug=True, stats=None)
        torch.cuda.synchronize()

        def run():
            MU.compute_fine_Rt_overlap(atten, inp["f_score"], inp["f_pts1"], inp["f_pts2"], None, cfg.dis_thres, stats=stats)

        for _ in range(3):
            run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run()
        # rotate nothing: atten (B x 16.8 MB) is larger than L2 at B = 16
        times = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / 20)
    res = dict(R=R, t=t, sc=sc, R2=R2, t2=t2, sc2=sc2)
    res.update({"d_" + k: v for k, v in dbg.items()})
    res.update({"e_" + k: v for k, v in dbg2.items()})
    torch.save({k: v.cpu() for k, v in res.items()}, out_path)
    print(json.dumps({"mode": os.environ.get("UPK_FINE_TMA", "1"), "fine_pose_us": [round(1e3 * x, 1) for x in times]}))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        return child(sys.argv[2])
    import torch
    outs = {}
    for mode in ("0", "1"):
        p = "/tmp/fine_tma_ab_%s.pt" % mode
        env = dict(os.environ, UPK_FINE_TMA=mode)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", p], env=env, capture_output=True, text=True)
        sys.stdout.write(r.stdout)
        if r.returncode != 0:
            sys.stderr.write(r.stderr[-4000:])
            raise SystemExit(1)
        outs[mode] = torch.load(p)
    bad = 0
    for k in outs["0"]:
        a, b = outs["0"][k], outs["1"][k]
        same = torch.equal(a, b)
        print("%-10s %s" % (k, "bit-identical" if same else "DIFFERS max|d| = %.3e" % (a.float() - b.float()).abs().max().item()))
        bad += 0 if same else 1
    print("A/B:", "OK" if bad == 0 else "%d tensors differ" % bad)
    raise SystemExit(0 if bad == 0 else 2)


if __name__ == "__main__":
    main()
