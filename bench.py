#!/usr/bin/env python
"""bench.py — instances posed / s (and hypotheses scored / s) of the correspondence-and-pose hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" = one pass of the hot path (unopose_b200.pipeline.run_hot_path) over one batch of B synthetic
instances per GPU (BASELINE.json configs[1] fine matching + configs[0] coarse solve, with the FPS /
ball-query / grouping stages of §8(a)).  Instances are independent, so ranks shard them with no
data-path collective ("scaling": "weak"); the only collective is the final gather of results.

Timing: W warm-up steps, then exactly K steps between (barrier + cuda synchronize), CUDA events on the
launching stream, MAX over ranks.  Inputs rotate over several resident sets whose total size exceeds
L2, so no step re-reads its inputs from L2.

`--impl reference`: the reference's CPU implementation of the same step (oracle port on the host
cores, all threads) on a bounded sample of the workload.
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
    ap.add_argument("--batch", type=int, default=16, help="instances per GPU per step (reference instance_batch_size=16)")
    ap.add_argument("--sets", type=int, default=4, help="resident input sets to rotate over (defeats L2 reuse)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=8, help="instances in the bounded CPU-baseline sample")
    ap.add_argument("--no-overlap", action="store_true", help="run the two chains of a step serially on one stream")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying CUDA graphs")
    ap.add_argument("--depth", type=int, default=2,
                    help="steps in flight: consecutive steps (different resident sets) replay on alternating "
                         "streams, so one step's serial FPS chain (16 SMs) overlaps the next step's full-GPU kernels")
    ap.add_argument("--no-zero-copy", action="store_true",
                    help="e2e: copy the full per-point feature tensors instead of gathering the FPS-selected rows "
                         "straight out of pinned host memory")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-torch-baseline", action="store_true")
    ap.add_argument("--no-widened", action="store_true", help="skip the informational timings of the f1/f2 kernels")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback (B200_PROFILING.md)")


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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s > 0.5 * max(sm)]
        return {"sm_mhz": statistics.median(busy or sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- reference (CPU) arm
def cpu_leg(cfg, sample_b, steps, warmup):
    """The reference's CPU path (oracle port) on `sample_b` instances per step; returns (inst/s, hyp/s, info)."""
    from oracle import hotpath_cpu
    from oracle import pose_oracle as PO
    from unopose_b200.pipeline import synthetic_inputs

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    inp = synthetic_inputs(1234, sample_b, cfg, device=None)
    for _ in range(warmup):
        hotpath_cpu.run_hot_path_cpu(inp, cfg, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        hotpath_cpu.run_hot_path_cpu(inp, cfg, threads)
    dt = (time.perf_counter() - t0) / steps
    # coarse solve alone (hypotheses scored / s)
    c_atten = PO.feature_similarity(inp["c_f1"], inp["c_f2"], "cosine", cfg.temp, True)
    t0 = time.perf_counter()
    for _ in range(max(steps, 2)):
        PO.coarse_pose(c_atten, inp["c_score"], inp["c_pts1"], inp["c_pts2"], None, cfg.n_proposal1, cfg.n_proposal2)
    dtc = (time.perf_counter() - t0) / max(steps, 2)
    return sample_b / dt, sample_b * cfg.n_proposal1 / dtc, dict(threads=threads, ms_per_step=dt * 1e3)


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    v, hyp, info = cpu_leg(cfg, args.cpu_sample, steps, warm)
    sample = "%d instances/step, %d timed steps, full hot path on host cores (torch CPU ops + C oracle)" % (
        args.cpu_sample, steps)
    line = {
        "impl": "reference", "metric": "instances_posed_per_s", "value": v, "unit": "instances/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": info["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, args.cpu_sample),
        "hypotheses_scored_per_s": hyp,
        "cpu_baseline": {"value": v, "unit": "instances/s", "cores": info["threads"], "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(cfg, batch):
    return {"workload": "coarse(196x196,C=256,H=%d,K=%d)+fine(2048x2048,C=256)+FPS(5000->2048,2048->196 x2)+"
                        "ball_query/group(r=.1/64,r=.2/256 x2 clouds)" % (cfg.n_proposal1, cfg.n_proposal2),
            "instances_per_gpu_per_step": batch, "l2": "inputs rotate over resident sets larger than L2",
            "parallelism": "instances sharded across ranks (no data-path collective)"}


# ----------------------------------------------------------------------------- our arm
def widened_leg(B, dev):
    """Device time of the kernels of the SURVEY.md §8 "next" rows at the real config (outside the headline step):
    f2 GeometricStructureEmbedding (N = 197, hidden 256, k = 3) and f1 PositionalEncoding (2048 points, two scales)."""
    import math

    import torch

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


def main():
    args = parse()
    from unopose_b200.pipeline import HotPathConfig

    cfg = HotPathConfig()
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
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    B = args.batch

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # resident input sets (each rank owns different instances: weak scaling)
    sets = [synthetic_inputs(1000 * rank + 17 * s, B, cfg, device=dev) for s in range(args.sets)]
    set_bytes = input_bytes(sets[0])
    results = []

    graphs = None if args.no_graph else [GraphedHotPath(s_, cfg, overlap=not args.no_overlap) for s_ in sets]

    depth = max(1, min(args.depth, len(sets)))
    if len(sets) % depth != 0:
        raise SystemExit("--sets must be a multiple of --depth (a resident set is always replayed on the same stream)")
    lanes = [torch.cuda.Stream(device=dev) for _ in range(depth)] if depth > 1 else None

    def step(i):
        # one CUDA graph per resident input set (static addresses); with depth > 1 step i runs on stream
        # i % depth, so `depth` consecutive steps (always different sets) are in flight at once
        def go():
            if graphs is not None:
                return graphs[i % len(sets)].replay()
            return run_hot_path(sets[i % len(sets)], cfg, overlap=not args.no_overlap)
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
        out = step(i)
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
        # the path's one collective: gather of per-instance results (R 9, t 3, score 1) on all ranks
        res = torch.cat([out["pred_R"].reshape(B, 9), out["pred_t"], out["pred_pose_score"].unsqueeze(1)], 1)
        gathered = [torch.empty_like(res) for _ in range(world)]
        dist.all_gather(gathered, res)
    ms_per_step = ms / args.steps
    value = B * world / (ms_per_step * 1e-3)

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
    e2e = {"value": B * world / (ms_e * 1e-3), "unit": "instances/s", "h2d_bytes_per_step": copied + pulled,
           "d2h_bytes_per_step": res_host.numel() * 4, "ms_per_step": ms_e,
           "h2d_copied_bytes": copied, "h2d_zero_copy_gather_bytes": pulled,
           "host_input_bytes": input_bytes(host_sets[0])}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- rooflines: one model per stage (algorithmic work per launch is defined in DESIGN.md §5);
    #      `roofline` is the model of the stage with the largest share of the (serial) step time
    pk = peaks()
    total_stage = sum(stage_ms.values())
    n1 = cfg.n_fine + 1
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    fp32_peak = sms * 128 * 1.965e9 / 1e12          # T lane-op/s, whole chip at max clock
    pe_bytes = sum((12.0 * 2 * cfg.n_fine + 4.0 * cfg.n_fine * ns + 16.0 * cfg.n_fine * ns) for _, ns in cfg.pe) * B
    fps_ops = lambda n, m: 10.0 * (m - 1) * n * B   # 10 lane-ops per distance update
    models = {
        "fine_similarity": ("k_similarity_tc2<stats> (tcgen05.mma.cta_group::2 3xTF32, CTA pairs on 256x256 tiles, background "
                            "row/column peeled, exponent sums of the assignment in the epilogue) + k_normalize_split x2",
                            "tensor", 2.0 * n1 * n1 * cfg.feat_dim * B),
        "fine_pose": ("k_fine_labels + k_fine_rows (2 reads of the 2049^2 fp32 matrix; pass 1 = exponent sums, fused "
                      "into the similarity GEMM's epilogue) + Kabsch + inliers", "hbm", 2.0 * n1 * n1 * 4 * B),
        "fps_template+gather": ("fps_kernel<512,10> (5000->2048; serial chain, one SM per instance)", "fp32",
                                fps_ops(cfg.n_template, cfg.n_fine)),
        "fps_sparse_ref+gather": ("fps_kernel<512,4> (2048->196)", "fp32", fps_ops(cfg.n_fine, cfg.n_coarse)),
        "fps_sparse_query+gather": ("fps_kernel<512,4> (2048->196)", "fp32", fps_ops(cfg.n_fine, cfg.n_coarse)),
        "ball_query+group_query": ("ball_group_kernel<2 scales> (fused ball query + grouping)", "hbm", pe_bytes),
        "ball_query+group_ref": ("ball_group_kernel<2 scales> (fused ball query + grouping)", "hbm", pe_bytes),
        "coarse_pose": ("k_score (K x 196 x 196 point pairs, 6 lane-ops each) + assignment/sampling/top-K", "fp32",
                        6.0 * cfg.n_proposal2 * cfg.n_coarse * cfg.n_coarse * B),
        "coarse_similarity": ("k_similarity_tc<0,3> (tcgen05 3xTF32, 197x197x256 per instance) + k_normalize_split x2",
                              "tensor", 2.0 * (cfg.n_coarse + 1) ** 2 * cfg.feat_dim * B),
    }

    def roof(stage):
        kname, bound, work = models[stage]
        dur_s = stage_ms[stage] * 1e-3
        extra = {}
        if bound == "tensor":
            achieved, peak, unit, bname = work / dur_s / 1e12, pk["tensor_sustained"], "TFLOP/s", "tensor"
            # the kernel issues 3 TF32 MMAs per product (3xTF32 split, fp32-level logits) and the TF32 rate of the
            # tensor pipe is half the bf16 rate `peak` was measured with
            extra = {"issued_tf32_tflops": 3.0 * achieved, "tf32_issue_peak": peak / 2.0,
                     "frac_of_tf32_issue_peak": 3.0 * achieved / (peak / 2.0),
                     "note": "achieved = ALGORITHMIC flops (2 n m c per instance) / stage duration incl. the operand "
                             "split kernels; peak = measured dense bf16"}
        elif bound == "fp32":
            achieved, peak, unit, bname = work / dur_s / 1e12, fp32_peak, "Tlane-op/s", "fp32-issue"
        else:
            achieved, peak, unit, bname = work / dur_s / 1e9, pk["hbm"], "GB/s", "hbm"
        return dict({"kernel": kname, "stage": stage, "stage_share_of_step": stage_ms[stage] / total_stage,
                     "bound": bname, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
                     "traffic": None, "peak_source": pk["source"] if bound != "fp32" else "#SM*128*1.965 GHz",
                     "duration_ms": stage_ms[stage]}, **extra)

    stage_rooflines = {k: roof(k) for k in stage_ms if k in models}
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath):       # measured DRAM traffic per launch from the committed ncu --set full capture
        tj = json.load(open(tpath))
        for k, r in stage_rooflines.items():
            if k in tj["stage_traffic_bytes"]:
                r["traffic"] = tj["stage_traffic_bytes"][k] * B / tj["batch"]
                r["traffic_source"] = "profiles/r1_traffic.json (ncu --set full, scaled by batch)"
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
        "stage_ms": stage_ms,
        "roofline": roofline,
        "stage_rooflines": stage_rooflines,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "resident_input_bytes": set_bytes * len(sets),
    }
    if world == 1 and not args.no_cpu_baseline:
        v, hyp, info = cpu_leg(cfg, args.cpu_sample, 3, 1)
        line["cpu_baseline"] = {"value": v, "unit": "instances/s", "cores": info["threads"], "kind": "port",
                                "sample": "%d instances/step x 3 steps of the same workload (oracle port: torch CPU "
                                          "ops + C restatement of the pointnet2 kernels)" % args.cpu_sample,
                                "hypotheses_scored_per_s": hyp}
    if world == 1 and not args.no_gpu_torch_baseline:
        try:
            line["gpu_torch_baseline"] = gpu_torch_leg(cfg, sets, B, dev)
        except Exception as ex:  # reported, never fatal
            line["gpu_torch_baseline"] = {"unavailable": repr(ex)[:200]}
    if world == 1 and not args.no_widened:
        try:
            line["widened_rows"] = widened_leg(B, dev)
        except Exception as ex:  # informational (SURVEY.md §8 "next" rows f1/f2), never fatal
            line["widened_rows"] = {"unavailable": repr(ex)[:200]}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def gpu_torch_leg(cfg, sets, B, dev):
    """The reference's GPU path on the same box and inputs: its torch ops (oracle port on CUDA tensors)
    plus its own pointnet2 extension compiled unmodified (oracle/_ref) — the ">= 10x" denominator of
    BASELINE.json.  Reported baseline only."""
    from oracle import pose_oracle as PO
    from oracle import ref_ext

    ext = ref_ext.load()
    if ext is None:
        return {"unavailable": "oracle/_ref/ref_pointnet2_ext.so not present"}

    def sample(pts, feats, m):
        idx = ext.furthest_point_sampling(pts.contiguous(), m)
        p = ext.gather_points(pts.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
        f = ext.gather_points(feats.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
        return p, f

    def ref_step(s):
        tem_sub, tem_f = sample(s["tem_pts"], s["tem_feats"], cfg.n_fine)
        sample(s["pts"], s["pts_feats"], cfg.n_coarse)
        sample(tem_sub, tem_f, cfg.n_coarse)
        ca = PO.feature_similarity(s["c_f1"], s["c_f2"], "cosine", cfg.temp, True)
        PO.coarse_pose(ca, s["c_score"], s["c_pts1"], s["c_pts2"], None, cfg.n_proposal1, cfg.n_proposal2)
        for cloud in (s["pts"], tem_sub):
            cloud = cloud.contiguous()
            cf = cloud.transpose(1, 2).contiguous()
            for r, ns in cfg.pe:
                ext.group_points(cf, ext.ball_query(cloud, cloud, r, ns))
        fa = PO.feature_similarity(s["f_f1"], s["f_f2"], "cosine", cfg.temp, True)
        return PO.fine_pose(fa, s["f_score"], s["f_pts1"], s["f_pts2"], None, cfg.dis_thres)

    for i in range(2):
        ref_step(sets[i % len(sets)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for i in range(n):
        ref_step(sets[i % len(sets)])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    # coarse solve alone
    ca = PO.feature_similarity(sets[0]["c_f1"], sets[0]["c_f2"], "cosine", cfg.temp, True)
    s = sets[0]
    e0.record()
    for i in range(n):
        PO.coarse_pose(ca, s["c_score"], s["c_pts1"], s["c_pts2"], None, cfg.n_proposal1, cfg.n_proposal2)
    e1.record()
    torch.cuda.synchronize()
    ms_c = e0.elapsed_time(e1) / n
    return {"value": B / (ms * 1e-3), "unit": "instances/s", "ms_per_step": ms, "kind": "port+reference_ext",
            "hypotheses_scored_per_s": B * cfg.n_proposal1 / (ms_c * 1e-3),
            "what": "reference torch-op sequence on CUDA tensors + reference pointnet2 _ext (unmodified, sm_100a)"}


if __name__ == "__main__":
    sys.exit(main())
