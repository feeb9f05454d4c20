#!/usr/bin/env python
"""bench.py — instances posed / s (and hypotheses scored / s) of the correspondence-and-pose hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference|reference-gpu]
    python bench.py --mode hyp-shard --H 5000 [--batch 8]          (BASELINE.json config 4)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" = one pass of the hot path (unopose_b200.pipeline.run_hot_path) over one batch of B synthetic instances per GPU
(BASELINE.json configs[1] fine matching + configs[0] coarse solve, with the FPS / ball-query / grouping stages of
SURVEY.md §8a).  Instances are independent, so ranks shard them ("scaling": "weak"); the path's one collective — the
all_gather of the per-instance result rows (R 9 | t 3 | score) — is issued EVERY step inside the timed region.

Timing: W warm-up steps, then exactly K steps between (barrier + cuda synchronize), CUDA events on the launching stream,
MAX over ranks.  Inputs rotate over several resident sets whose total size exceeds L2, so no step re-reads its inputs
from L2.

`--impl reference`     the reference's CPU implementation of the same step on the host cores (its own staged Python
                       from baseline/_ref when present, else the oracle port), same B / steps / warm-up.
`--impl reference-gpu` the UNMODIFIED reference on this GPU (its torch ops + its own `_ext` compiled for sm_100a).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=None, help="instances per GPU per step (default 16 = the reference's "
                                                            "instance_batch_size; 8 in --mode hyp-shard)")
    ap.add_argument("--sets", type=int, default=None,
                    help="resident input sets to rotate over (defeats L2 reuse; default 2 x --depth)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--mode", default="step", choices=["step", "hyp-shard"])
    ap.add_argument("--H", type=int, nargs="*", default=None, help="hyp-shard: hypotheses per instance (default sweep)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                    help="hyp-shard: candidate / score exchange through peer memory (fused into the kernels) or NCCL")
    ap.add_argument("--no-overlap", action="store_true", help="run the two chains of a step serially on one stream")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying CUDA graphs")
    ap.add_argument("--depth", type=int, default=3,
                    help="steps in flight: consecutive steps (different resident sets) replay on alternating "
                         "streams, so one step's serial FPS chain (16 SMs) overlaps the next step's full-GPU kernels")
    ap.add_argument("--no-zero-copy", action="store_true",
                    help="e2e: copy the full per-point feature tensors instead of gathering the FPS-selected rows "
                         "straight out of pinned host memory")
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of back-to-back steps for the sustained figure (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-widened", action="store_true", help="skip the informational timings of the f1/f2 kernels")
    ap.add_argument("--no-peaks", action="store_true", help="skip the live TF32 dense-peak measurement")
    ap.add_argument("--no-numa", action="store_true", help="do not pin the process to the GPU's NUMA node")
    args = ap.parse_args()
    if args.sets is None:
        args.sets = 2 * max(1, args.depth)
    return args


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def measure_tf32_peak(dev, seconds=1.0):
    """Dense TF32 tensor peak measured the way MEASURED_PEAKS.json measures bf16: torch.matmul 8192^3 (cuBLAS, TF32
    allowed), best of 10 (burst) and back to back for `seconds` (sustained)."""
    n = 8192
    a = torch.randn(n, n, device=dev)
    b = torch.randn(n, n, device=dev)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for _ in range(2):
            a @ b
        torch.cuda.synchronize(dev)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        reps = max(4, int(seconds * 1e3 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            a @ b
        e1.record()
        torch.cuda.synchronize(dev)
        sus = e0.elapsed_time(e1) / reps
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    fl = 2.0 * n ** 3
    return {"tf32_tflops": fl / best * 1e-9, "tf32_tflops_sustained": fl / sus * 1e-9,
            "how": "torch.matmul fp32 8192^3 with allow_tf32 (2 N^3): best of 10 and back to back for %.1f s" % seconds}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s > 0.5 * max(sm)]
        return {"sm_mhz": statistics.median(busy or sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


def pin_to_gpu_numa_node(local):
    """Bind this process (and therefore its pinned host allocations, first touch) to the NUMA node of its GPU."""
    try:
        bdf = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bdf.count(":") == 2 and len(bdf.split(":")[0]) == 8:    # 00000000:1b:00.0 -> 0000:1b:00.0 (sysfs)
            bdf = bdf[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return {"numa_node": None, "pinned": False, "why": "single NUMA node"}
        cpus = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        ids = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-")
                ids |= set(range(int(a), int(b) + 1))
            elif part:
                ids.add(int(part))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
        return {"numa_node": node, "pinned": bool(ids), "cpus": len(ids)}
    except Exception as ex:  # noqa: BLE001
        return {"numa_node": None, "pinned": False, "why": repr(ex)[:120]}


# ----------------------------------------------------------------------------- reference arms
def _staged_reference(need_ext):
    try:
        from baseline import refgpu

        if refgpu.available():
            return refgpu.load(need_ext=need_ext)
    except Exception:  # noqa: BLE001
        pass
    return None


def cpu_leg(cfg, B, steps, warmup):
    """The reference's CPU path on B instances per step: its own staged Python (baseline/_ref) when present — its three
    CUDA-only pointnet2 ops served by the C restatement — else the oracle port.  -> (inst/s, hyp/s, info)."""
    from oracle import hotpath_cpu
    from oracle import pose_oracle as PO
    from unopose_b200.pipeline import synthetic_inputs

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    inp = synthetic_inputs(1234, B, cfg, device=None)
    ns = _staged_reference(need_ext=False)
    if ns is not None:
        from baseline.ref_hot_path import ref_step_cpu

        step = lambda: ref_step_cpu(ns, inp, cfg, threads)  # noqa: E731
        coarse = lambda a: ns.model_utils.compute_coarse_Rt_overlap(a, inp["c_score"], inp["c_pts1"], inp["c_pts2"], None,  # noqa: E731
                                                                    cfg.n_proposal1, cfg.n_proposal2)
        kind = "reference"
    else:
        step = lambda: hotpath_cpu.run_hot_path_cpu(inp, cfg, threads)  # noqa: E731
        coarse = lambda a: PO.coarse_pose(a, inp["c_score"], inp["c_pts1"], inp["c_pts2"], None, cfg.n_proposal1,  # noqa: E731
                                          cfg.n_proposal2)
        kind = "port"
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    c_atten = PO.feature_similarity(inp["c_f1"], inp["c_f2"], "cosine", cfg.temp, True)
    nc = max(min(steps, 5), 2)
    t0 = time.perf_counter()
    for _ in range(nc):
        coarse(c_atten)
    dtc = (time.perf_counter() - t0) / nc
    return B / dt, B * cfg.n_proposal1 / dtc, dict(threads=threads, ms_per_step=dt * 1e3, kind=kind)


def gpu_reference_leg(cfg, sets, B, dev, steps, warmup):
    """The UNMODIFIED reference on this GPU: its own model_utils / pointnet2_utils (baseline/_ref) and its own `_ext`
    compiled for sm_100a — the ">= 10x" denominator of BASELINE.json, same inputs, B, steps and warm-up, CUDA events."""
    ns = _staged_reference(need_ext=True)
    if ns is None:
        return {"unavailable": "reference not staged under baseline/_ref or oracle/_ref/ref_pointnet2_ext.so missing"}
    from baseline.ref_hot_path import ref_step

    for i in range(max(warmup, 1)):
        ref_step(ns, sets[i % len(sets)], cfg)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        ref_step(ns, sets[i % len(sets)], cfg)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    s = sets[0]
    ca = ns.model_utils.compute_feature_similarity(s["c_f1"], s["c_f2"], "cosine", cfg.temp, True)
    ns.model_utils.compute_coarse_Rt_overlap(ca, s["c_score"], s["c_pts1"], s["c_pts2"], None, cfg.n_proposal1, cfg.n_proposal2)
    torch.cuda.synchronize(dev)
    e0.record()
    for i in range(steps):
        ns.model_utils.compute_coarse_Rt_overlap(ca, s["c_score"], s["c_pts1"], s["c_pts2"], None, cfg.n_proposal1,
                                                 cfg.n_proposal2)
    e1.record()
    torch.cuda.synchronize(dev)
    ms_c = e0.elapsed_time(e1) / steps
    return {"value": B / (ms * 1e-3), "unit": "instances/s", "ms_per_step": ms, "kind": "reference", "steps": steps,
            "warmup": warmup, "instances_per_step": B,
            "hypotheses_scored_per_s": B * cfg.n_proposal1 / (ms_c * 1e-3), "coarse_solve_ms": ms_c,
            "what": "unmodified reference (baseline/_ref: core.unopose.utils.model_utils + pointnet2_utils, its `_ext` "
                    "built unmodified for sm_100a), eager PyTorch on the same GPU, resident inputs, CUDA events"}


def run_reference(args, cfg):
    """`--impl reference`: rank 0 alone times the reference's CPU path, same B / steps / warm-up / config as our arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    B = args.batch
    v, hyp, info = cpu_leg(cfg, B, args.steps, args.warmup)
    what = ("the reference's own Python (baseline/_ref) on CPU tensors, its CUDA-only pointnet2 ops served by the C "
            "restatement oracle/pointnet2_oracle.c" if info["kind"] == "reference" else
            "oracle port (torch CPU ops + C restatement of the pointnet2 kernels)")
    sample = "%d instances/step x %d timed steps of the same workload on all host threads: %s" % (B, args.steps, what)
    line = {
        "impl": "reference", "metric": "instances_posed_per_s", "value": v, "unit": "instances/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": info["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(cfg, B), steps_in_flight=max(1, min(args.depth, args.sets))),
        "hypotheses_scored_per_s": hyp,
        "cpu_baseline": {"value": v, "unit": "instances/s", "cores": info["threads"], "kind": info["kind"], "sample": sample},
        "e2e": {"value": v, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(cfg, batch):
    return {"workload": "coarse(196x196,C=256,H=%d,K=%d)+fine(2048x2048,C=256)+FPS(5000->2048,2048->196 x2)+"
                        "ball_query/group(r=.1/64,r=.2/256 x2 clouds)" % (cfg.n_proposal1, cfg.n_proposal2),
            "instances_per_gpu_per_step": batch, "l2": "inputs rotate over resident sets larger than L2",
            "parallelism": "instances sharded across ranks; per-step all_gather of the result rows"}


# ----------------------------------------------------------------------------- informational legs
def widened_leg(B, dev):
    """Device time of the kernels of the SURVEY.md §8 "next" rows at the real config (outside the headline step):
    f2 GeometricStructureEmbedding (N = 197, hidden 256, k = 3) and f1 PositionalEncoding (2048 points, two scales)."""
    import math

    from unopose_b200.modules import geo
    from unopose_b200.modules.matching import PositionalEncoding

    def timeit(fn, it=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / it

    g = torch.Generator(device="cpu").manual_seed(5)
    N, C, k = 197, 256, 3
    p = torch.randn(B, N - 1, 3, generator=g)
    pts = torch.cat([torch.ones(B, 1, 3), p / p.norm(dim=2).max(dim=1)[0].view(B, 1, 1)], 1).to(dev)
    dterm = torch.exp(torch.arange(0, C, 2).float() * (-math.log(10000.0) / C)).to(dev)
    w = [((torch.rand(*s, generator=g) * 2 - 1) / math.sqrt(C)).to(dev) for s in ((C, C), (C,), (C, C), (C,))]
    fa = 180.0 / (15 * math.pi)
    t_geo = timeit(lambda: geo.geometric_embedding(pts, dterm, w[0], w[1], w[2], w[3], 0.2, fa, k))
    flops = 2.0 * B * N * N * (1 + k) * C * C
    pe = PositionalEncoding(256, r1=0.1, r2=0.2, nsample1=64, nsample2=256, use_lrf=True, use_xyz=True).to(dev).eval()
    cloud = torch.randn(B, 2048, 3, generator=g)
    cloud = (cloud / cloud.norm(dim=2).max(dim=1)[0].view(B, 1, 1)).to(dev)
    with torch.no_grad():
        t_pe = timeit(lambda: pe(cloud), it=3, warm=1)
    return {"geometric_embedding": {"ms_per_call": t_geo, "clouds_per_call": B, "algorithmic_tflops": flops / t_geo * 1e-9,
                                    "issued_tf32_tflops": 3 * flops / t_geo * 1e-9 * 1275.0 / 1213.0},
            "positional_encoding": {"ms_per_call": t_pe, "clouds_per_call": B,
                                    "what": "fused ball query/grouping + k_lrf_group + k_shared_mlp_max x2 + Conv1d (cuBLAS)"}}


def h2d_ceiling(host_sets, fed, dev, barrier, steps):
    """The copies of an e2e step alone (same pinned buffers, same bytes, no compute), all ranks at once: the PCIe /
    host-memory ceiling the end-to-end figure can be held against."""
    src = host_sets[0]
    dst = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in src.items() if not k.startswith("_")}
    for k, v in dst.items():
        v.copy_(src[k], non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        s = host_sets[i % len(host_sets)]
        for k, v in dst.items():
            if k not in fed.zero_copy:
                v.copy_(s[k], non_blocking=True)
    e1.record()
    torch.cuda.synchronize(dev)
    barrier()
    nbytes = sum(v.numel() * v.element_size() for k, v in dst.items() if k not in fed.zero_copy)
    return e0.elapsed_time(e1) / steps, nbytes


# ----------------------------------------------------------------------------- config 4: hypothesis sharding
def run_hyp_shard(args, cfg):
    """BASELINE.json config 4: hypothesis-count sweep, the pool of ONE instance batch split across the N ranks, both
    collectives (all_gather of the candidate lists, all_reduce(MAX) of the score table) inside the timed loop; beside it
    the single-GPU solver (upk_coarse_pose) on the same inputs."""
    from unopose_b200 import _lib
    from unopose_b200 import model_utils as MU
    from unopose_b200.dist import HypothesisShardedCoarse
    from unopose_b200.synthetic import matching_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    B = args.batch or 8
    K = cfg.n_proposal2
    Hs = args.H or [1000, 2000, 5000, 10000, 20000, 50000, 100000]
    n = cfg.n_coarse

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    d = matching_batch(4242, B, n, cfg.feat_dim)             # identical on every rank
    T = {k: torch.from_numpy(v).to(dev) for k, v in d.items() if k in ("pts1", "pts2", "f1", "f2", "score")}
    atten = MU.compute_feature_similarity(T["f1"], T["f2"], "cosine", cfg.temp, True)
    rows = []
    for H in Hs:
        u = torch.rand(B, 3 * H, generator=torch.Generator().manual_seed(H)).to(dev)
        solver = HypothesisShardedCoarse(B, n, n, H, K, dev, exchange=args.exchange)   # auto: peer memory when mappable
        solver_nccl = HypothesisShardedCoarse(B, n, n, H, K, dev, exchange="nccl") if solver.exchange == "p2p" else None
        single = lambda: MU._coarse(atten, T["score"], T["pts1"], T["pts2"], None, H, K, u=u)  # noqa: E731
        shard = lambda: solver.run(atten, T["score"], T["pts1"], T["pts2"], u)  # noqa: E731
        shard_nccl = lambda: solver_nccl.run(atten, T["score"], T["pts1"], T["pts2"], u)  # noqa: E731
        R1, t1, s1 = single()
        R2, t2, s2, _ = shard()
        torch.cuda.synchronize()
        same = bool(torch.equal(R1, R2) and torch.equal(t1, t2) and torch.equal(s1, s2))
        if solver_nccl is not None:
            R3, t3, s3, _ = shard_nccl()
            torch.cuda.synchronize()
            same = same and bool(torch.equal(R1, R3) and torch.equal(t1, t3) and torch.equal(s1, s3))
        res = {"exchange": solver.exchange}
        variants = [("sharded", shard)] + ([("sharded_nccl", shard_nccl)] if solver_nccl is not None else []) + [("single_gpu", single)]
        for name, fn in variants:
            graph = None
            if not args.no_graph:
                try:                                           # NCCL collectives are graph-capturable
                    side = torch.cuda.Stream(device=dev)
                    side.wait_stream(torch.cuda.current_stream(dev))
                    with torch.cuda.stream(side):
                        for _ in range(2):
                            fn()
                    torch.cuda.current_stream(dev).wait_stream(side)
                    torch.cuda.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        fn()
                except Exception:  # noqa: BLE001
                    graph = None
                    torch.cuda.synchronize()
            run = graph.replay if graph is not None else fn
            for _ in range(max(args.warmup, 3)):
                run()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                run()
            e1.record()
            torch.cuda.synchronize()
            barrier()
            ms = e0.elapsed_time(e1) / args.steps
            if dist is not None:
                tt = torch.tensor([ms], device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ms = float(tt.item())
            res[name] = {"ms_per_solve": ms, "hypotheses_per_s": B * H / (ms * 1e-3), "cuda_graph": graph is not None}
            del graph
        # the two collectives alone (same buffers), to name the limiter
        coll = {}
        if dist is not None:
            sv = solver_nccl or solver
            for cname, fn in (("all_gather_candidates", lambda: dist.all_gather_into_tensor(sv.allc, sv.cand)),
                              ("all_reduce_max_scores", lambda: dist.all_reduce(sv.scores, op=dist.ReduceOp.MAX))):
                for _ in range(5):
                    fn()
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(50):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                coll[cname + "_us"] = e0.elapsed_time(e1) / 50 * 1e3
            coll["all_gather_bytes_per_rank"] = solver.cand.numel() * 4
            coll["all_reduce_bytes"] = solver.scores.numel() * 4
            if solver.px is not None:
                coll["peer_exchange_timed_out"] = solver.px.timed_out()
        rows.append(dict(H=H, K=K, B=B, bit_identical_to_single_gpu=same, **res, collectives=coll))
    if rank == 0:
        best = max(rows, key=lambda r: r["sharded"]["hypotheses_per_s"])
        line = {"metric": "hypotheses_scored_per_s", "value": best["sharded"]["hypotheses_per_s"], "unit": "hypotheses/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": best["sharded"]["ms_per_solve"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "mode": "hyp-shard",
                "config": {"workload": "coarse solve 196x196, K=%d, hypothesis pool of one %d-instance batch split over %d rank(s); "
                                       "value = the best H of the sweep" % (K, B, world), "H_of_value": best["H"],
                           "l2": "coarse working set (0.5 MB / instance) is L2-resident by nature"},
                "exchange": rows[0].get("exchange"),
                "sweep": rows, "gpu_launches": int(args.steps * (8 if rows[0].get("exchange") == "p2p" else 10))}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


# ----------------------------------------------------------------------------- our arm
def main():
    args = parse()
    from unopose_b200.pipeline import HotPathConfig

    cfg = HotPathConfig()
    if args.mode == "hyp-shard":
        return run_hyp_shard(args, cfg)
    if args.batch is None:
        args.batch = 16
    if args.impl == "reference":
        return run_reference(args, cfg)

    from unopose_b200 import _lib
    from unopose_b200 import model_utils as MU
    from unopose_b200.pipeline import GraphedHotPath, HostFedHotPath, input_bytes, run_hot_path, synthetic_inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = {"pinned": False} if args.no_numa else pin_to_gpu_numa_node(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    B = args.batch

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # resident input sets (each rank owns different instances: weak scaling)
    sets = [synthetic_inputs(1000 * rank + 17 * s, B, cfg, device=dev) for s in range(args.sets)]
    if args.impl == "reference-gpu":
        if rank == 0:
            r = gpu_reference_leg(cfg, sets, B, dev, args.steps, max(args.warmup, 1))
            line = {"impl": "reference-gpu", "metric": "instances_posed_per_s", "value": r.get("value"), "unit": "instances/s",
                    "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": r.get("ms_per_step"),
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": workload_config(cfg, B), "gpu_reference": r, "gpu_launches": 0}
            print(json.dumps(line))
        if dist is not None:
            dist.destroy_process_group()
        return 0
    _lib.load()
    set_bytes = input_bytes(sets[0])

    graphs = None if args.no_graph else [GraphedHotPath(s_, cfg, overlap=not args.no_overlap) for s_ in sets]

    depth = max(1, min(args.depth, len(sets)))
    if len(sets) % depth != 0:
        raise SystemExit("--sets must be a multiple of --depth (a resident set is always replayed on the same stream)")
    lanes = [torch.cuda.Stream(device=dev) for _ in range(depth)] if depth > 1 else None
    # the path's one collective, every step: all_gather of this step's (B, 13) result rows; NCCL runs it on its own
    # stream behind the step that produced the rows, so it overlaps the next step
    gathered = [torch.empty((world * B, 13), dtype=torch.float32, device=dev) for _ in range(len(sets))] if dist is not None else None
    pending = []
    last_work = {}

    def step(i):
        # one CUDA graph per resident input set (static addresses); with depth > 1 step i runs on stream
        # i % depth, so `depth` consecutive steps (always different sets) are in flight at once
        def go():
            k = i % len(sets)
            if k in last_work:
                last_work.pop(k).wait()      # the gather that still reads this set's result rows (stream-side wait)
            if graphs is not None:
                o = graphs[k].replay()
            else:
                o = run_hot_path(sets[k], cfg, overlap=not args.no_overlap)
            if dist is not None:
                last_work[k] = dist.all_gather_into_tensor(gathered[k], o["result"], async_op=True)
                pending.append(last_work[k])
            return o
        if lanes is None:
            return go()
        with torch.cuda.stream(lanes[i % depth]):
            return go()

    def fork():
        if lanes is not None:
            cur = torch.cuda.current_stream(dev)
            for l_ in lanes:
                l_.wait_stream(cur)

    def join():
        for w_ in pending:
            w_.wait()
        pending.clear()
        if lanes is not None:
            cur = torch.cuda.current_stream(dev)
            for l_ in lanes:
                cur.wait_stream(l_)

    # allocator / module-load settling (untimed, not counted as warm-up): every resident set is
    # seen twice so the two stream pools of the caching allocator reach steady state
    fork()
    for i in range(2 * len(sets)):
        step(i)
    for i in range(max(args.warmup, 3)):
        step(i)
    join()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.1)
    # launches per step are counted on an eager (non-graph) step: a graph replay re-issues the same kernels
    torch.cuda.synchronize()
    lc0 = _lib.launch_count()
    run_hot_path(sets[0], cfg, overlap=not args.no_overlap)
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - lc0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    e0.record()
    step_ev[0].record()
    fork()
    for i in range(args.steps):
        step(i)
        step_ev[i + 1].record(lanes[i % depth] if lanes is not None else None)
    join()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    # completion-to-completion intervals (with depth > 1 steps overlap, so these are not step latencies)
    done_t = [step_ev[0].elapsed_time(step_ev[i + 1]) for i in range(args.steps)]
    per_step = [b_ - a_ for a_, b_ in zip([0.0] + sorted(done_t)[:-1], sorted(done_t))]
    launches = launches_per_step * args.steps
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = B * world / (ms_per_step * 1e-3)

    # ---- the same loop held for seconds: does the figure survive sustained load (clocks, power)?
    sustained = None
    if args.sustain > 0:
        n_sus = max(args.steps, int(args.sustain * 1e3 / ms_per_step))
        s2 = ClockSampler(local)
        barrier()
        if rank == 0:
            s2.start()
        e0.record()
        fork()
        for i in range(n_sus):
            step(i)
        join()
        e1.record()
        torch.cuda.synchronize()
        barrier()
        ms_s = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms_s], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_s = float(t.item())
        sustained = {"seconds": ms_s * 1e-3, "steps": n_sus, "ms_per_step": ms_s / n_sus,
                     "value": B * world / (ms_s / n_sus * 1e-3), "unit": "instances/s",
                     "clocks": s2.stop() if rank == 0 else None}

    # ---- coarse solve alone: hypotheses scored / s (a2..a6 on resident atten)
    c_att = [MU.compute_feature_similarity(s["c_f1"], s["c_f2"], "cosine", cfg.temp, True) for s in sets]
    torch.cuda.synchronize()

    def coarse_only(i):
        s = sets[i % len(sets)]
        MU.compute_coarse_Rt_overlap(c_att[i % len(sets)], s["c_score"], s["c_pts1"], s["c_pts2"], None,
                                     cfg.n_proposal1, cfg.n_proposal2)

    for i in range(3):
        coarse_only(i)
    barrier()
    e0.record()
    for i in range(args.steps):
        coarse_only(i)
    e1.record()
    torch.cuda.synchronize()
    ms_c = e0.elapsed_time(e1) / args.steps
    if dist is not None:
        t = torch.tensor([ms_c], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_c = float(t.item())
    hyp_per_s = B * world * cfg.n_proposal1 / (ms_c * 1e-3)

    # ---- per-stage device times: every stage is captured into its own CUDA graph and replayed back to back
    #      (pure device time on one stream, no Python launch overhead between the events); rank 0 reports
    stage_ms = {}
    plan = []
    run_hot_path(sets[0], cfg, stages=plan)
    cap = torch.cuda.Stream(device=dev)
    for name, fn in plan:
        cap.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cap):
            fn()                                   # eager once: later stages read this stage's outputs
        torch.cuda.current_stream(dev).wait_stream(cap)
        torch.cuda.synchronize()
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_):
            fn()
        reps = 10
        g_.replay()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(reps):
            g_.replay()
        s1.record()
        torch.cuda.synchronize()
        stage_ms[name] = s0.elapsed_time(s1) / reps
        del g_

    # ---- end to end through the public API with HOST (pinned) buffers: H2D of every step input,
    #      D2H of the step result, both inside the timed region
    host_sets = [synthetic_inputs(5000 + 1000 * rank + 17 * s, B, cfg, pin=True) for s in range(2)]
    fed = HostFedHotPath(cfg, B, dev, overlap=not args.no_overlap, use_graph=not args.no_graph,
                         zero_copy=not args.no_zero_copy)

    def e2e_step(i):
        # prefetch the NEXT step's inputs (copy stream) while this step computes; every timed step
        # therefore contains one full H2D of a step's inputs, one compute and one D2H of the result
        fed.stage((i + 1) % 2, host_sets[(i + 1) % len(host_sets)])
        return fed.run(i % 2)

    fed.stage(0, host_sets[0])
    for i in range(3):
        e2e_step(i)
    barrier()
    e0.record()
    for i in range(3, 3 + args.steps):
        e2e_step(i)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms_e = e0.elapsed_time(e1) / args.steps
    if dist is not None:
        t = torch.tensor([ms_e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e = float(t.item())
    res_host = fed.result
    copied, pulled = fed.pcie_bytes_per_step(host_sets[0])
    ms_copy, copy_bytes = h2d_ceiling(host_sets, fed, dev, barrier, args.steps)
    if dist is not None:
        t = torch.tensor([ms_copy], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_copy = float(t.item())
    e2e = {"value": B * world / (ms_e * 1e-3), "unit": "instances/s", "h2d_bytes_per_step": copied + pulled,
           "d2h_bytes_per_step": res_host.numel() * 4, "ms_per_step": ms_e,
           "h2d_copied_bytes": copied, "h2d_zero_copy_gather_bytes": pulled,
           "host_input_bytes": input_bytes(host_sets[0]),
           "h2d_gbs_per_gpu": (copied + pulled) / (ms_e * 1e-3) / 1e9,
           "h2d_gbs_all_gpus": world * (copied + pulled) / (ms_e * 1e-3) / 1e9,
           "copies_alone": {"ms_per_step": ms_copy, "bytes": copy_bytes, "gbs_per_gpu": copy_bytes / (ms_copy * 1e-3) / 1e9,
                            "gbs_all_gpus": world * copy_bytes / (ms_copy * 1e-3) / 1e9,
                            "what": "the cudaMemcpyAsync part of a step alone, all ranks at once: host-memory / PCIe ceiling"},
           "numa": numa}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- rooflines: one model per stage (algorithmic work per launch is defined in DESIGN.md §5);
    #      `roofline` is the model of the stage with the largest share of the (serial) step time.  Stages are timed in
    #      isolation (10 graph replays, ~ms): the BURST peaks apply; the sustained figures are listed beside them.
    pk = peaks()
    tf32 = None
    if world == 1 and not args.no_peaks:
        try:
            tf32 = measure_tf32_peak(dev)
        except Exception as ex:  # noqa: BLE001
            tf32 = {"unavailable": repr(ex)[:120]}
    total_stage = sum(stage_ms.values())
    n1 = cfg.n_fine + 1
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    fp32_peak = sms * 128 * 1.965e9 / 1e12          # T lane-op/s, whole chip at max clock
    pe_bytes = sum((12.0 * 2 * cfg.n_fine + 4.0 * cfg.n_fine * ns + 16.0 * cfg.n_fine * ns) for _, ns in cfg.pe) * B
    fps_ops = lambda n, m: 10.0 * (m - 1) * n * B   # 10 lane-ops per distance update
    fine_bytes = B * (2.0 * n1 * cfg.feat_dim * 4 + 2.0 * cfg.n_fine * 12 + 2.0 * cfg.n_fine * 4 + 1.0 * n1 * n1 * 4)
    models = {
        "fine_similarity": ("k_similarity_tc2<stats, fp16> (tcgen05.mma.cta_group::2 kind::f16, 3xFP16 split, CTA pairs on "
                            "256x256 tiles, background row/column peeled, exponent sums of the assignment in the epilogue) "
                            "+ k_normalize_split x2", "tensor-f16", 2.0 * n1 * n1 * cfg.feat_dim * B),
        "fine_pose": ("k_fine_labels_tma + k_fine_rows_tma (2 reads of the 2049^2 fp32 logits through a TMA box ring; "
                      "pass 1 = exponent sums, fused into the similarity GEMM's epilogue) + merges + Kabsch + inliers",
                      "hbm", 2.0 * n1 * n1 * 4 * B),
        "fps_template+gather": ("fps_kernel<512,10> (5000->2048; serial chain, one SM per instance)", "fp32",
                                fps_ops(cfg.n_template, cfg.n_fine)),
        "fps_sparse_ref+gather": ("fps_kernel<512,4> (2048->196)", "fp32", fps_ops(cfg.n_fine, cfg.n_coarse)),
        "fps_sparse_query+gather": ("fps_kernel<512,4> (2048->196)", "fp32", fps_ops(cfg.n_fine, cfg.n_coarse)),
        "ball_query+group_query": ("ball_group_kernel<2 scales> (fused ball query + grouping)", "hbm", pe_bytes),
        "ball_query+group_ref": ("ball_group_kernel<2 scales> (fused ball query + grouping)", "hbm", pe_bytes),
        "coarse_pose": ("k_score (K x 196 x 196 point pairs, 6 lane-ops each) + k_coarse_assign_exact + sampling / top-K",
                        "fp32", 6.0 * cfg.n_proposal2 * cfg.n_coarse * cfg.n_coarse * B),
        "coarse_similarity": ("k_similarity_tc<0,3> (tcgen05 3xTF32, 197x197x256 per instance) + k_normalize_split x2",
                              "tensor-tf32", 2.0 * (cfg.n_coarse + 1) ** 2 * cfg.feat_dim * B),
    }

    def roof(stage):
        kname, bound, work = models[stage]
        dur_s = stage_ms[stage] * 1e-3
        extra = {}
        if bound.startswith("tensor"):
            achieved, unit, bname = work / dur_s / 1e12, "TFLOP/s", "tensor"
            if bound == "tensor-f16":
                peak, psrc = pk["tensor_burst"], pk["source"] + " dense bf16/fp16, burst (the stage is timed in isolation)"
                extra = {"peak_sustained": pk["tensor_sustained"]}
            else:
                have = tf32 is not None and "tf32_tflops" in tf32
                peak = tf32["tf32_tflops"] if have else pk["tensor_burst"] / 2.0
                psrc = ("measured in this run: " + tf32["how"] + ", burst") if have else "dense bf16 burst / 2 (TF32 peak not measured)"
                if have:
                    extra = {"peak_sustained": tf32["tf32_tflops_sustained"]}
            # three MMAs per product (hi*hi + hi*lo + lo*hi, fp32-level logits): issued = 3 x algorithmic
            extra.update({"issued_tflops": 3.0 * achieved, "frac_issued": 3.0 * achieved / peak,
                          "note": "achieved = ALGORITHMIC flops (2 n m c per instance) / stage duration incl. the "
                                  "operand split kernels; the kernel issues 3 tensor-core products per algorithmic one"})
        elif bound == "fp32":
            achieved, peak, unit, bname, psrc = work / dur_s / 1e12, fp32_peak, "Tlane-op/s", "fp32-issue", "#SM*128*1.965 GHz"
        else:
            achieved, peak, unit, bname, psrc = work / dur_s / 1e9, pk["hbm"], "GB/s", "hbm", pk["source"]
        return dict({"kernel": kname, "stage": stage, "stage_share_of_step": stage_ms[stage] / total_stage,
                     "bound": bname, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
                     "traffic": None, "peak_source": psrc, "duration_ms": stage_ms[stage]}, **extra)

    stage_rooflines = {k: roof(k) for k in stage_ms if k in models}
    for tname in ("r2_traffic.json", "r1_traffic.json"):   # measured DRAM traffic per launch from the committed ncu --set full capture
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            for k, r in stage_rooflines.items():
                if k in tj["stage_traffic_bytes"]:
                    r["traffic"] = tj["stage_traffic_bytes"][k] * B / tj["batch"]
                    r["traffic_source"] = "profiles/%s (ncu --set full, scaled by batch)" % tname
            break
    # the fine stage as a whole against HBM: algorithmic bytes = operands + ONE write of atten (the API returns it)
    fine_ms = stage_ms.get("fine_similarity", 0.0) + stage_ms.get("fine_pose", 0.0)
    fine_stage = {"algorithmic_bytes": fine_bytes, "ms": fine_ms, "achieved_gbs": fine_bytes / (fine_ms * 1e-3) / 1e9 if fine_ms else None,
                  "frac_of_hbm": fine_bytes / (fine_ms * 1e-3) / 1e9 / pk["hbm"] if fine_ms else None}
    # the dominant kernel = largest share of the GPU's capacity (duration x fraction of the SMs it occupies): the FPS
    # kernels run one CTA per instance (B of the SMs), concurrently with the full-GPU kernels of the step in flight
    sm_share = {k: (min(1.0, B / sms) if k.startswith("fps_") else 1.0) for k in stage_ms}
    for k, r in stage_rooflines.items():
        r["sm_share"] = sm_share[k]
    dom = max(stage_rooflines, key=lambda k: stage_ms[k] * sm_share[k])
    roofline = stage_rooflines[dom]

    line = {
        "metric": "instances_posed_per_s", "value": value, "unit": "instances/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(cfg, B), steps_in_flight=depth),
        "hypotheses_scored_per_s": hyp_per_s, "coarse_solve_ms": ms_c,
        "step_ms_min_median_max": [min(per_step), statistics.median(per_step), max(per_step)],
        "sustained": sustained,
        "stage_ms": stage_ms,
        "roofline": roofline,
        "stage_rooflines": stage_rooflines,
        "fine_stage_vs_hbm": fine_stage,
        "measured_tf32_peak": tf32,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "collective": ({"op": "all_gather_into_tensor of the (B,13) result rows, every step, inside the timed region",
                        "bytes_per_rank": B * 13 * 4} if dist is not None else None),
        "clocks": clocks,
        "resident_input_bytes": set_bytes * len(sets),
    }
    if world == 1 and not args.no_cpu_baseline:
        v, hyp, info = cpu_leg(cfg, B, 3, 1)
        what = ("the reference's own Python (baseline/_ref) on CPU tensors; its CUDA-only pointnet2 ops served by the C "
                "restatement" if info["kind"] == "reference" else "oracle port: torch CPU ops + C restatement of the pointnet2 kernels")
        line["cpu_baseline"] = {"value": v, "unit": "instances/s", "cores": info["threads"], "kind": info["kind"],
                                "sample": "%d instances/step x 3 steps of the same workload (%s)" % (B, what),
                                "hypotheses_scored_per_s": hyp}
    if world == 1 and not args.no_gpu_reference:
        try:
            line["gpu_reference"] = gpu_reference_leg(cfg, sets, B, dev, args.steps, max(args.warmup, 3))
            if "value" in line["gpu_reference"]:
                line["speedup_vs_gpu_reference"] = {"device_resident": value / line["gpu_reference"]["value"],
                                                    "e2e_vs_resident_reference": e2e["value"] / line["gpu_reference"]["value"],
                                                    "hypotheses_per_s": hyp_per_s / line["gpu_reference"]["hypotheses_scored_per_s"]}
        except Exception as ex:  # reported, never fatal
            line["gpu_reference"] = {"unavailable": repr(ex)[:200]}
    if world == 1 and not args.no_widened:
        try:
            line["widened_rows"] = widened_leg(B, dev)
        except Exception as ex:  # informational (SURVEY.md §8 "next" rows f1/f2), never fatal
            line["widened_rows"] = {"unavailable": repr(ex)[:200]}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
