#!/usr/bin/env python
"""Throughput benchmark of the association hot path (driver contract: one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload = BASELINE.json configs[3] / SURVEY §8d cfg4, the configuration the metric is quoted on:
pp_pv_40e_dualadd_subabs_C (Fusion C, minus_abs, dual_add), N = M = 128 detections per frame,
P = 512 LiDAR points per detection, 64x64 crops.  A "step" = one pass of forward + association LP
over `--pairs` frame-pairs per GPU (weak scaling: per-GPU work fixed).  Frame-pairs are independent
units, sharded over ranks with no data-path collective; the only collective is the final gather of
the assignment indices, inside the timed region.

value : frame-pairs/s with inputs resident in HBM (CUDA events, max over ranks).
e2e   : same metric through the public API with HOST (pinned) inputs: H2D of crops/points and D2H
        of the assignment results inside the timed region.
roofline: dominant kernel = tma::gemm_tma_kernel in conv mode, the TMA-fed tcgen05 3x3-conv contraction of the
        VGG trunk (12 launches per chunk, 83 % of the algorithmic FLOPs, ~50 % of the step), timed per launch
        with CUDA events on the launching stream (library hook mmmot_timing_*).
cpu_baseline / --impl reference: the oracle port of the reference's PyTorch-CPU path (the reference
        is pure Python and /root/reference does not exist on the GPU box) on all host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(fusion="C", affinity_op="minus_abs", softmax_mode="dual_add", neg_threshold=0.2, n=128, pts=512, hw=64)
METRIC = "frame-pairs/sec at N=128 dets"
# algorithmic FLOPs per frame-pair at cfg4: VGG 641.4 G, affinity 83.76 G, PointNet 48.8 G, total 775.7 G
FLOP_PER_PAIR = 775.7e9


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = max([int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = sorted(float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit())
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm),
                "power_w": pw[len(pw) // 2] if pw else None}


def pick_cpu_threads():
    """torch's intra-op scaling on many-core hosts is poor for these shapes (128 threads were 14x slower
    than 8 on the GPU box); time one small frame-pair at a few thread counts and keep the fastest."""
    from mmmot_b200.synthetic import synthetic_pair, synthetic_state_dict
    from oracle import torch_ref
    ncpu = os.cpu_count() or 1
    sd = synthetic_state_dict(CFG["fusion"], seed=0)
    dets, info, split = synthetic_pair(16, 16, 64, CFG["hw"], seed=0)
    best, best_t = 1, float("inf")
    for th in sorted({min(ncpu, t) for t in (8, 16, 32, 64, ncpu)}):
        torch.set_num_threads(th)
        torch_ref.forward(sd, dets, info, split, CFG["fusion"], CFG["affinity_op"], CFG["softmax_mode"], CFG["neg_threshold"])
        t = time.perf_counter()
        torch_ref.forward(sd, dets, info, split, CFG["fusion"], CFG["affinity_op"], CFG["softmax_mode"], CFG["neg_threshold"])
        dt = time.perf_counter() - t
        if dt < best_t:
            best, best_t = th, dt
    return best


def oracle_pairs_per_s(n_pairs, threads, budget_s=150.0):
    """The reference's CPU path (oracle port of TrackingNet.forward + HiGHS restatement of the LP).
    Bounded sample: stops early once `budget_s` seconds of CPU work are spent (>= 1 pair is always timed)."""
    from mmmot_b200.synthetic import synthetic_pair, synthetic_state_dict
    from oracle import lp_ref, torch_ref
    torch.set_num_threads(threads)
    sd = synthetic_state_dict(CFG["fusion"], seed=0)
    t_tot = 0.0
    for p in range(n_pairs):
        dets, info, split = synthetic_pair(CFG["n"], CFG["n"], CFG["pts"], CFG["hw"], seed=p)
        t = time.perf_counter()
        det, link, new, end, _ = torch_ref.forward(sd, dets, info, split, CFG["fusion"], CFG["affinity_op"],
                                                   CFG["softmax_mode"], CFG["neg_threshold"])
        lp_ref.milp_solve(det[2], [link[0][2:3]], new[2], end[2], [CFG["n"], CFG["n"]])
        t_tot += time.perf_counter() - t
        if t_tot > budget_s:
            n_pairs = p + 1
            break
    return n_pairs / t_tot, t_tot


def config_dict(pairs, world):
    return {"workload": "cfg4 pp_pv_40e_dualadd_subabs_C: Fusion C, minus_abs, dual_add, N=M=128 dets/frame, "
                        "P=512 LiDAR pts/det, 64x64 crops; forward + association LP",
            "pairs_per_gpu_per_step": pairs, "global_pairs_per_step": pairs * world,
            "parallelism": f"frame-pair sharding x{world}, final gather only",
            "l2_policy": "inputs larger than L2 (%.1f GB per GPU per step vs 126 MB L2)" % (pairs * 14.16e6 / 1e9)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = pick_cpu_threads()                          # fastest thread count on this host (<= all cores)
    t0 = time.perf_counter()
    rate, secs = oracle_pairs_per_s(max(args.steps, 1), threads)    # one frame-pair per "step"
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "frame-pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config_dict(1, 1), note="CPU: one frame-pair per step (bounded sample of the same workload)"),
            "cpu_baseline": {"value": rate, "unit": "frame-pairs/s", "cores": threads, "host_cores": os.cpu_count(), "kind": "port",
                             "sample": f"{max(args.steps, 1)} frame-pairs of the cfg4 shape, oracle port of the reference "
                                       "PyTorch-CPU forward + HiGHS MILP restatement of ortools_solve (OR-tools absent)"},
            "e2e": {"value": rate, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--pairs", type=int, default=int(os.environ.get("MMMOT_BENCH_PAIRS", "128")),
                    help="frame-pairs per GPU per step")
    ap.add_argument("--cpu-pairs", type=int, default=2, help="frame-pairs timed for cpu_baseline (rank 0)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--kseg", type=int, default=-1, help="tcgen05 conv K-segment length in 32-chunks (0 = off; default: library default)")
    ap.add_argument("--engine", default="auto", choices=["auto", "fp32", "tcgen05"],
                    help="contraction engine (A/B runs; default auto = tcgen05 for this workload)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import mmmot_b200
    from mmmot_b200 import _lib
    from mmmot_b200.synthetic import synthetic_state_dict
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    mmmot_b200.set_engine(args.engine)
    if args.kseg >= 0:
        lib.mmmot_set_kseg(args.kseg)

    n, pts, hw = CFG["n"], CFG["pts"], CFG["hw"]
    L, B = 2 * n, args.pairs
    net = mmmot_b200.TrackingNet(2, appear_skippool=True, score_arch="branch_cls", score_fusion_arch=CFG["fusion"],
                                 affinity_op=CFG["affinity_op"], softmax_mode=CFG["softmax_mode"],
                                 neg_threshold=CFG["neg_threshold"], test_mode=2, dropblock=0)
    net.load_state_dict(synthetic_state_dict(CFG["fusion"], seed=0))
    net.cuda(dev).eval()

    # synthetic inputs of the cfg4 shape, generated on the device (seeded per rank), mirrored to pinned host memory
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    crops = torch.randn(B * L, 3, hw, hw, device=dev, generator=g)
    centre = torch.rand(B * L, 1, 3, device=dev, generator=g) * torch.tensor([60.0, 40.0, 2.0], device=dev) + \
        torch.tensor([0.0, -20.0, -2.0], device=dev)
    points = (torch.randn(B * L, pts, 3, device=dev, generator=g) * torch.tensor([2.0, 1.0, 0.8], device=dev) + centre).reshape(-1, 3)
    split = torch.arange(0, B * L * pts + 1, pts, dtype=torch.int32)
    h_crops = torch.empty(crops.shape, dtype=torch.float32, pin_memory=True).copy_(crops)
    h_points = torch.empty(points.shape, dtype=torch.float32, pin_memory=True).copy_(points)
    h_match = torch.empty(B, n, dtype=torch.int32, pin_memory=True)
    h_flags = torch.empty(3, B, L, dtype=torch.float32, pin_memory=True)
    d_crops2, d_points2 = torch.empty_like(crops), torch.empty_like(points)

    def step_resident():
        o = net.predict_batch(crops, points, split, n)
        if world > 1:
            gathered = [torch.empty_like(o["match"]) for _ in range(world)]
            dist.all_gather(gathered, o["match"])
        return o

    # e2e: pinned host -> device copies are pipelined against compute in sub-batches (copy stream + events);
    # every byte of every step's inputs crosses PCIe inside the timed region, results come back D2H.
    nsub = 4 if B % 4 == 0 and B >= 8 else 1
    sb = B // nsub
    copy_stream = torch.cuda.Stream(device=dev)
    ev_copied = [torch.cuda.Event() for _ in range(nsub)]
    ev_used = [torch.cuda.Event() for _ in range(nsub)]
    sub_split = torch.arange(0, sb * L * pts + 1, pts, dtype=torch.int32)
    for e in ev_used:
        e.record()

    def step_e2e():
        cur = torch.cuda.current_stream(dev)
        for i in range(nsub):
            c0, c1 = i * sb * L, (i + 1) * sb * L
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_used[i])             # previous step finished reading this slice
                d_crops2[c0:c1].copy_(h_crops[c0:c1], non_blocking=True)
                d_points2[c0 * pts:c1 * pts].copy_(h_points[c0 * pts:c1 * pts], non_blocking=True)
                ev_copied[i].record(copy_stream)
        outs = []
        for i in range(nsub):
            c0, c1 = i * sb * L, (i + 1) * sb * L
            cur.wait_event(ev_copied[i])
            o = net.predict_batch(d_crops2[c0:c1], d_points2[c0 * pts:c1 * pts], sub_split, n)
            ev_used[i].record(cur)
            p0, p1 = i * sb, (i + 1) * sb
            h_match[p0:p1].copy_(o["match"], non_blocking=True)
            h_flags[0, p0:p1].copy_(o["assign_det"], non_blocking=True)
            h_flags[1, p0:p1].copy_(o["assign_new"], non_blocking=True)
            h_flags[2, p0:p1].copy_(o["assign_end"], non_blocking=True)
            outs.append(o["match"])
        if world > 1:
            allm = torch.cat(outs, 0)
            gathered = [torch.empty_like(allm) for _ in range(world)]
            dist.all_gather(gathered, allm)
        return outs

    def timed(fn, steps, warmup, with_hooks=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        l0 = lib.mmmot_launch_count()
        if with_hooks:
            lib.mmmot_timing_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        hook = None
        if with_hooks:
            lib.mmmot_timing_enable(0)
            tm, fl, cnt = ctypes.c_double(), ctypes.c_double(), ctypes.c_long()
            lib.mmmot_timing_collect(ctypes.byref(tm), ctypes.byref(fl), ctypes.byref(cnt))
            hook = (tm.value, fl.value, cnt.value)
        launches = lib.mmmot_launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, launches, hook

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches, hook = timed(step_resident, args.steps, args.warmup, with_hooks=True)
    clocks = sampler.summary() if sampler else None
    sampler2 = ClockSampler(local) if rank == 0 else None
    if sampler2:
        sampler2.start()
    ms_e2e, _, _ = timed(step_e2e, args.steps, max(args.warmup, 1))
    clocks_e2e = sampler2.summary() if sampler2 else None

    if rank == 0:
        peaks, how = load_peaks()
        value = B * world * args.steps / (ms / 1e3)
        e2e = B * world * args.steps / (ms_e2e / 1e3)
        conv_ms, conv_flop, conv_n = hook
        achieved = conv_flop / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        tc_engine = args.engine != "fp32"
        traffic = None
        try:    # dram bytes per frame-pair of the 12 conv launches, from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "r01_conv_traffic.json")) as f:
                tr = json.load(f)
            if tc_engine:
                traffic = tr["dram_bytes_per_pair"] * (B / max(conv_n / (12 * args.steps), 1)) / 12
        except Exception:
            pass
        roofline = {"bound": "tensor",
                    "kernel": ("tma::gemm_tma_kernel / gemm_tma_px_kernel, conv mode (TMA-fed tcgen05 3x3-conv contraction of the VGG "
                               "trunk, layers 1..12, FP16 hi/lo split: 3 MMAs per algorithmic MAC)") if tc_engine else
                              "gemm_simt_kernel<XM_CONV3> (VGG 3x3 conv contraction, FP32 FFMA engine)",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "mma_issue_tflops": 3 * achieved if tc_engine else None,
                    "mma_issue_frac": 3 * achieved / peak if tc_engine else None,
                    "peak_source": f"{how} bf16_tflops_sustained (kernel timed inside a long step); 'achieved' counts "
                                   "ALGORITHMIC FLOPs (2*Cout*9Cin*pixels per launch); the tensor pipe executes 3x that",
                    "launches_timed": conv_n, "avg_launch_ms": conv_ms / max(conv_n, 1),
                    "share_of_step": conv_ms / ms, "traffic": traffic,
                    "traffic_note": "avg dram bytes per launch, scaled from profiles/r01_conv_traffic.json (ncu --set full)"}
        cpu = None
        if not args.no_cpu and world == 1:      # the CPU baseline is reported at N=1 only
            threads = pick_cpu_threads()
            rate, secs = oracle_pairs_per_s(args.cpu_pairs, threads)
            cpu = {"value": rate, "unit": "frame-pairs/s", "cores": threads, "host_cores": os.cpu_count(), "kind": "port",
                   "sample": f"{args.cpu_pairs} frame-pairs of the same cfg4 workload ({secs:.1f} s), oracle port of the reference "
                             "PyTorch-CPU forward + HiGHS restatement of the LP"}
        h2d = h_crops.numel() * 4 + h_points.numel() * 4 + split.numel() * 4
        d2h = h_match.numel() * 4 + h_flags.numel() * 4
        line = {"metric": METRIC, "value": value, "unit": "frame-pairs/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32 (tcgen05 engine: FP16 hi/lo split operands, fp32 accumulate)" if args.engine != "fp32" else "f32", "data": "synthetic", "config": dict(config_dict(B, world), engine=args.engine),
                "clocks": clocks, "clocks_e2e": clocks_e2e,
                "e2e": {"value": e2e, "unit": "frame-pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches),
                "algorithmic_tflops": value * FLOP_PER_PAIR / 1e12 / world,
                "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
