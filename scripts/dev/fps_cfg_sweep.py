"""FPS thread-count configurations (UPK_FPS_CFG): time 5000 -> 2048 and 2048 -> 196 at B = 16 and check the indices
against the default configuration bit for bit (one process per configuration: the knob is read once)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child(out):
    import torch
    from unopose_b200.pointnet2 import _ext as mine
    from util_clouds import batch_clouds
    dev = torch.device("cuda:0")
    res = {}
    for name, n, m, kind in (("tem", 5000, 2048, "surface"), ("sparse", 2048, 196, "surface"), ("quant", 5000, 2048, "quantised")):
        pts = torch.from_numpy(batch_clouds(3, 16, n, "surface")).to(dev)
        if kind == "quantised":
            pts = (pts * 16).round() / 16     # coarse grid: many duplicate points and exact distance ties
        idx = mine.furthest_point_sampling(pts, m)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            mine.furthest_point_sampling(pts, m)
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        res[name] = idx.cpu()
        print("cfg %s %-6s %d->%d  %.1f us" % (os.environ.get("UPK_FPS_CFG", "0"), name, n, m, min(ts)))
    torch.save(res, out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2])
        raise SystemExit(0)
    import torch
    base = None
    for cfg in ("0", "5", "6"):
        p = "/tmp/fps_cfg_%s.pt" % cfg
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", p], env=dict(os.environ, UPK_FPS_CFG=cfg),
                           capture_output=True, text=True)
        sys.stdout.write(r.stdout)
        if r.returncode:
            sys.stderr.write(r.stderr[-3000:])
            raise SystemExit(1)
        d = torch.load(p)
        if base is None:
            base = d
        else:
            print("cfg", cfg, {k: bool(torch.equal(v, base[k])) for k, v in d.items()})
