#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric: M points/s of Waymo-shape hard voxelization
(config C4: 64 frames x 180 000 points x 5 features, voxel [0.1,0.1,0.15], range
[-75.2,-75.2,-2,75.2,75.2,4], max_points 5, max_voxels 150 000) on N B200s, next to the HBM
roofline and the reference's CPU op timed on the same host.

    python bench.py --gpus 1 --steps K --warmup W                 (our arm, one GPU)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N GPUs)
    python bench.py --impl reference --gpus N --steps K --warmup W    (reference CPU arm)

A step = one pass of the hot path over the config's batch of 64 synthetic frames.  Frames shard
across GPUs with no collective: STRONG scaling (contiguous blocks of 64 / N frames
per GPU; `--scaling weak` gives every rank its own 64 frames).  The only communication is a
barrier and a MAX over the ranks' device times, on the host (gloo): no NCCL on this path.
Prints ONE JSON line.
"""
import argparse
import glob
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = "C4"
KEPT_POINTS_PER_FRAME = {"C4": 156_000}  # points surviving the max_points cap (oracle, seed 4000)
METRIC = "hard_voxelize_throughput"
UNIT = "Mpoints/s"
# ProfScope name of the library (pcfe_profile_report) -> kernel name in an ncu report
SCOPE_TO_KERNEL = {"memset_ctl": ["hvb_zero_kernel"], "hvb_bin": ["hvb_bin_kernel"],
                   "hvb_bucket": ["hvb_bucket_rec_kernel", "hvb_bucket_rank_kernel", "hvb_bucket_small_kernel"],
                   "hvb_scan_firsts": ["hvb_scan_firsts_kernel"], "hv_scan_flags": ["hv_scan_flags_kernel"],
                   "hvb_order": ["hvb_order_kernel"],
                   "hvb_expand": ["hvb_expand_rec_kernel", "hvb_expand_words_kernel", "hvb_expand_pipe_kernel",
                                  "hvb_expand_fixed_kernel", "hvb_expand_kernel"],
                   "hv_slow_fallback": ["hvg_slow_frame_kernel"], "hvc_group": ["hvc_group_kernel"]}


def ncu_traffic(workload, scopes):
    """DRAM bytes per step from the newest committed ncu capture of this workload
    (profiles/*_traffic_<workload>.json, written by tests/native/ncu_traffic.py from an
    `ncu --set full` report).  Returns (bytes or None, per-kernel dict, source): None when the
    capture's kernel list does not cover the kernels this run launched -- a stale capture must not
    pass for a measurement."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"*_traffic_{workload}.json")))
    if not files:
        return None, {}, "no profiles/*_traffic_%s.json" % workload
    d = json.load(open(files[-1]))
    per = d.get("kernels", {})
    src = os.path.relpath(files[-1], ROOT) + " <- " + d.get("source", "?")
    tot, missing = 0, []
    for sc in scopes:
        hit = [k for k in SCOPE_TO_KERNEL.get(sc, []) if k in per]
        if hit:
            tot += sum(per[k]["dram_read"] + per[k]["dram_write"] for k in hit)
        elif sc not in ("hv_slow_fallback", "memset_ctl"):  # (no-op / counter-zeroing launches: no DRAM traffic)
            missing.append(sc)
    if missing:
        return None, per, src + " (STALE: no capture of " + ", ".join(missing) + ")"
    per = dict(per, _frames_per_step=int(d.get("frames_per_step", 64)))
    return int(tot), per, src


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=0, help="frames per step over all GPUs (default: the config's 64)")
    ap.add_argument("--order", default="shuffled", choices=["shuffled", "sweep"],
                    help="point order of the synthetic frames: PointShuffle'd (the configs) or a spinning sensor's firing order")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the step's frames are split over the GPUs (default); weak: every GPU gets all of them")
    ap.add_argument("--workload", default=WORKLOAD, choices=["C1", "C4", "C5"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-chunk", type=int, default=8, help="frames per pipelined chunk of the e2e measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the packed / fused-encoder timings (\"fused\" key)")
    ap.add_argument("--frames-in-flight", type=int, default=0)
    ap.add_argument("--hv-wave", type=int, default=0, help="tuning: frames per wave (0 = library default)")
    ap.add_argument("--hv-bucket-avg", type=int, default=0, help="tuning: target points per bucket")
    ap.add_argument("--debug", action="append", default=[], metavar="NAME=VALUE",
                    help="tuning: pcfe_debug_set knob (repeatable), e.g. --debug mega_d1=3")
    ap.add_argument("--ref-procs", type=int, default=0, help="reference arm: worker processes (default: all cores, <= 64)")
    ap.add_argument("--sample-frames", type=int, default=0,
                    help="reference arm: frames per step (default: the whole batch; the cpu_baseline leg of our line uses 32)")
    return ap.parse_args()


def workload_config(args):
    from detmatch_b200 import synth
    cfg = dict(synth.CONFIGS[args.workload])
    if args.frames > 0:
        cfg["frames"] = args.frames
    elif args.workload == "C1":
        cfg["frames"] = 16  # the config is a single frame; 16 is the batch the survey profiles it on
    cfg["index"] = int(args.workload[1])
    return cfg


def algorithmic_bytes(n, c, p, m_list):
    """every input row read once, every returned output element written once:
    N*C*4 + M*(P*C*4 + 3*4 + 4) per frame, M = returned voxel_num."""
    return sum(n * c * 4 + m * (p * c * 4 + 16) for m in m_list)


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        return self

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU op (oracle/_ref) on the host cores
# --------------------------------------------------------------------------------------------
def _ref_worker(conn, cfg, frame_ids, use_ref):
    """One single-threaded worker: owns a block of the step's frames, voxelizes them with the
    reference op on demand."""
    import torch
    torch.set_num_threads(1)
    from detmatch_b200 import synth
    if use_ref:
        from oracle import ref
        ref.module()
    else:
        from oracle import oracle
        oracle.lib()
    frames = [synth.lidar_frame(cfg["n"], cfg["c"], synth.seed_for(cfg["index"], fid), cfg["r_max"]) for fid in frame_ids]
    conn.send("ready")
    while True:
        cmd = conn.recv()
        if cmd != "run":
            break
        m = 0
        for pts in frames:
            if use_ref:
                # voxelize.py:46-58 call pattern, including the three new_zeros
                v, c, n = ref.voxelization(pts, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], cfg["max_voxels"])
            else:
                v, c, n = oracle.hard_voxelize(pts.numpy(), cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], cfg["max_voxels"])
            m += int(n.shape[0])
        conn.send(m)


def run_reference(args, quiet=False, sample_frames=0):
    """Times the reference's CPU hard_voxelize on this host: one single-threaded worker process per
    core; a step is the SAME batch our arm voxelizes (all of the config's frames, split evenly over
    the workers) unless `sample_frames` bounds it (the cpu_baseline leg of our own line)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    import multiprocessing as mp
    cfg = workload_config(args)
    from oracle import ref
    use_ref = ref.available()
    if use_ref:
        ref.module()  # the parent loads oracle/_ref too: the workers are forked children
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    frames_total = sample_frames if sample_frames > 0 else cfg["frames"]
    procs = args.ref_procs if args.ref_procs > 0 else min(cores, 64)
    procs = max(1, min(procs, frames_total))
    blocks = [list(range(k * frames_total // procs, (k + 1) * frames_total // procs)) for k in range(procs)]
    ctx = mp.get_context("fork")
    workers = []
    for k in range(procs):
        parent, child = ctx.Pipe()
        p = ctx.Process(target=_ref_worker, args=(child, cfg, blocks[k], use_ref), daemon=True)
        p.start()
        workers.append((p, parent))
    for _, conn in workers:
        assert conn.recv() == "ready"

    def step():
        t0 = time.perf_counter()
        for _, conn in workers:
            conn.send("run")
        ms = [conn.recv() for _, conn in workers]
        return time.perf_counter() - t0, ms

    for _ in range(max(args.warmup, 1)):
        step()
    times, ms = [], None
    for _ in range(max(args.steps, 1)):
        dt, ms = step()
        times.append(dt)
    for p, conn in workers:
        conn.send("stop")
        p.join(timeout=5)
    t_step = sum(times) / len(times)
    pts_per_step = frames_total * cfg["n"]
    value = pts_per_step / t_step / 1e6
    kind = "reference" if use_ref else "port"
    sample = (f"{frames_total} frames/step ({cfg['n']} pts x {cfg['c']}) over {procs} single-threaded worker processes "
              f"(one per core), {len(times)} steps; "
              f"{'oracle/_ref = reference voxelization_cpu.cpp compiled in place' if use_ref else 'oracle C port'}")
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": max(args.warmup, 1), "ms_per_step": round(t_step * 1e3, 3), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": dict(_config_dict(cfg, args, frames_per_step=frames_total),
                       frames_per_gpu_per_step=len(frame_block(frames_total, 0, max(args.gpus, 1)))
                       if args.scaling == "strong" else frames_total),
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": procs, "kind": kind, "sample": sample,
                         "host_cores_visible": cores, "mean_voxels_per_frame": round(sum(ms) / frames_total, 1)},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not quiet:
        print(json.dumps(line), flush=True)
    return line


def _config_dict(cfg, args, frames_per_step):
    return {"workload": f"{args.workload}: Waymo-shape hard voxelization" if args.workload == "C4" else args.workload,
            "frames_per_step": frames_per_step, "points_per_frame": cfg["n"], "features": cfg["c"],
            "voxel_size": cfg["voxel_size"], "point_cloud_range": cfg["point_cloud_range"],
            "max_num_points": cfg["max_num_points"], "max_voxels": cfg["max_voxels"],
            "generator": "LiDAR-like (SURVEY 8(d)), seed = 1000*config + frame" + ("" if args.order == "shuffled" else ", sweep order"),
            "l2": "inputs+outputs per step (~0.94 GB at 64 C4 frames) exceed the 126 MB L2; no explicit flush",
            "sharding": "frames split over the GPUs in contiguous blocks, no collective"}


# --------------------------------------------------------------------------------------------
# multi-GPU plumbing (no data-path collective: frames shard, only times are reduced, on the host)
# --------------------------------------------------------------------------------------------
def frame_block(total_frames, rank, world):
    """Global frame ids owned by `rank` under strong scaling: contiguous blocks, disjoint, balanced."""
    return list(range(rank * total_frames // world, (rank + 1) * total_frames // world))


def frame_seeds(cfg_index, rank, frames_per_rank):
    """Weak scaling: every rank its own `frames_per_rank` frames (contiguous, disjoint across ranks)."""
    return [rank * frames_per_rank + k for k in range(frames_per_rank)]


def max_over_ranks(value, device=None):
    """MAX of a python float over all ranks (identity when not distributed), reduced on the host."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from detmatch_b200 import _cabi, synth
    from detmatch_b200.ops.voxel import HardVoxelizeBatchPlan

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # host-side process group: the path has no exchange step, only the ranks' times are reduced
        dist.init_process_group("gloo")
    cfg = workload_config(args)
    N, C = cfg["n"], cfg["c"]
    P, V = cfg["max_num_points"], cfg["max_voxels"]
    F_total = cfg["frames"]
    if args.scaling == "strong":
        frame_ids = frame_block(F_total, rank, world)
        total_frames = F_total
    else:
        frame_ids = frame_seeds(cfg["index"], rank, F_total)
        total_frames = F_total * world
    F = len(frame_ids)
    if args.hv_wave:
        _cabi.debug_set("hv_wave", args.hv_wave)
    if args.hv_bucket_avg:
        _cabi.debug_set("hv_bucket_avg", args.hv_bucket_avg)
    for kv in args.debug:
        name, _, val = kv.partition("=")
        _cabi.debug_set(name, int(val))

    # synthetic frames, generated on the host
    host = [synth.lidar_frame(N, C, synth.seed_for(cfg["index"], fid), cfg["r_max"], order=args.order).pin_memory()
            for fid in frame_ids]
    pts = [h.to(dev, non_blocking=True) for h in host]
    torch.cuda.synchronize(dev)
    plan = HardVoxelizeBatchPlan([N] * F, C, cfg["voxel_size"], cfg["point_cloud_range"], P, V, dev,
                                 frames_in_flight=args.frames_in_flight).bind(pts)
    L = _cabi.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # clocks are sampled from before the warm-up (nvidia-smi needs ~0.2 s to start) to the end of
    # the timed region; everything in between runs the same kernels
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        plan.run()
    barrier()
    m_list = plan.voxel_num.cpu().tolist()

    # ---- timed region: K steps, device time, inputs resident in HBM -------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = L.pcfe_launch_count()
    barrier()
    e0.record()
    for _ in range(args.steps):
        plan.run()
    e1.record()
    barrier()
    launches = L.pcfe_launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    if rank == 0 and clocks and clocks["samples"] < 3:
        # timed region shorter than the sampling period: sample an identical untimed loop
        sampler = ClockSampler(local_rank).start()
        t_end = time.time() + 0.6
        while time.time() < t_end:
            plan.run()
        torch.cuda.synchronize(dev)
        clocks = sampler.stop()
        clocks["note"] = "timed region < sampling period; sampled during an identical untimed loop right after"
    ms_step = max_over_ranks(ms_total) / args.steps
    total_points = total_frames * N
    value = total_points / (ms_step * 1e-3) / 1e6

    # ---- per-kernel attribution (separate, untimed pass) -------------------------------------
    kernels = None
    if rank == 0:
        _cabi.profile(True)
        for _ in range(3):
            plan.run()
        torch.cuda.synchronize(dev)
        rep = _cabi.profile_report()
        _cabi.profile(False)
        tot = sum(v[0] for v in rep.values()) or 1.0
        kernels = {k: {"ms_per_step": round(v[0] / 3, 4), "launches_per_step": v[1] // 3, "share": round(v[0] / tot, 3)}
                   for k, v in rep.items()}

    # ---- end to end through the package's host-buffer API -----------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, cfg, host, dev, total_frames, barrier)

    # ---- the other scaling mode, a few steps (extra key, not the headline) -------------------
    weak = None
    if world > 1 and args.scaling == "strong" and not args.no_extras:
        weak = run_weak_extra(cfg, rank, world, dev, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline ----------------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    algo = algorithmic_bytes(N, C, P, m_list)  # this rank's share of the step
    achieved = algo / (ms_step * 1e-3) / 1e9
    traffic, per_kernel, traffic_src = ncu_traffic(args.workload, list(kernels or {}))
    cap_frames = per_kernel.get("_frames_per_step", 64)
    if traffic is not None and F != cap_frames:
        traffic = int(traffic * F / cap_frames)  # scaled from the captured step's frame count
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic,
                "kernel": "hard-voxelize launch sequence per GPU (" + ", ".join(k for k in (kernels or {})) + "); "
                          "achieved = algorithmic bytes of the rank's step / CUDA-event step time",
                "algorithmic_bytes_per_step": algo, "peak_source": peak_src + " (of measured)",
                "traffic_source": "sum of dram__bytes_read+write over the step's kernels, ncu --set full: " + traffic_src,
                "mean_voxels_per_frame": round(sum(m_list) / max(len(m_list), 1), 1)}
    if kernels and "hvb_expand" in kernels:
        # dominant kernel (the expansion): writes every returned element once, reads the kept rows and
        # one first-point index per voxel (the records of multi-point voxels are not counted)
        kept = KEPT_POINTS_PER_FRAME.get(args.workload)
        m_tot = sum(m_list)
        k_bytes = m_tot * (P * C * 4 + 16) + m_tot * 4 + (kept * F * C * 4 if kept else 0)
        k_ms = kernels["hvb_expand"]["ms_per_step"]
        kt = next((per_kernel[k] for k in SCOPE_TO_KERNEL["hvb_expand"] if k in per_kernel), None)
        roofline["dominant_kernel"] = {"name": "hvb_expand", "ms_per_launch": k_ms,
                                       "algorithmic_bytes_per_launch": k_bytes,
                                       "achieved": round(k_bytes / (k_ms * 1e-3) / 1e9, 1),
                                       "frac": round(k_bytes / (k_ms * 1e-3) / 1e9 / peak, 4),
                                       "traffic": int((kt["dram_read"] + kt["dram_write"]) * F / cap_frames) if kt else None}

    fused = None
    if not args.no_extras and P == 5 and C in (4, 5):
        fused = run_fused_extras(cfg, pts, plan, dev)
    latency = None
    if not args.no_extras and world == 1:
        latency = run_latency_extras(cfg, pts, dev)

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        cpu_baseline = cpu_baseline_subprocess(args)
    reference_cuda = None
    if not args.no_cpu_baseline and not args.no_extras and world == 1:
        reference_cuda = run_reference_cuda(cfg, pts, plan, m_list, dev)

    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "ours",
        "config": dict(_config_dict(cfg, args, frames_per_step=total_frames), frames_per_gpu_per_step=F),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks, "kernels": kernels, "fused": fused, "latency": latency, "weak": weak,
        "reference_cuda": reference_cuda,
        "comm": "gloo (host): barrier + MAX of the ranks' device times; no data-path collective, no NCCL" if world > 1 else None,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_weak_extra(cfg, rank, world, dev, barrier, steps=20):
    """Weak-scaling figure next to the strong-scaling headline: every rank voxelizes its own full
    batch (the config's 64 frames).  Device time, MAX over ranks."""
    import torch
    from detmatch_b200 import synth
    from detmatch_b200.ops.voxel import HardVoxelizeBatchPlan
    F, N, C = cfg["frames"], cfg["n"], cfg["c"]
    pts = [synth.lidar_frame(N, C, synth.seed_for(cfg["index"], fid), cfg["r_max"]).to(dev)
           for fid in frame_seeds(cfg["index"], rank, F)]
    plan = HardVoxelizeBatchPlan([N] * F, C, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"],
                                 cfg["max_voxels"], dev).bind(pts)
    for _ in range(5):
        plan.run()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        plan.run()
    b.record()
    barrier()
    ms = max_over_ranks(a.elapsed_time(b)) / steps
    return {"scaling": "weak", "frames_per_gpu_per_step": F, "ms_per_step": round(ms, 4),
            "value": round(world * F * N / (ms * 1e-3) / 1e6, 1), "unit": UNIT, "steps": steps}


def run_reference_cuda(cfg, pts, plan, m_list, dev, frames=2):
    """Third column (optional): the reference's OWN CUDA op -- hard_voxelize_gpu,
    voxelization_cuda.cu:184-326, compiled unmodified for sm_100a into oracle/_ref/detmatch_ref_cuda.so --
    per frame on the same GPU, called as the reference's wrapper calls it (three new_zeros + the op,
    voxelize.py:46-58).  A baseline next to the CPU arm, not a product path; its output is also compared
    with ours on the frames it runs."""
    import torch
    try:
        from oracle import ref
        if not ref.cuda_available():
            return {"unavailable": "oracle/_ref/detmatch_ref_cuda.so not built"}
        ref.cuda_module()
        P, V = cfg["max_num_points"], cfg["max_voxels"]
        frames = min(frames, len(pts))
        ref.cuda_voxelization(pts[0], cfg["voxel_size"], cfg["point_cloud_range"], P, V)  # warm-up
        torch.cuda.synchronize(dev)
        same = True
        t0 = time.perf_counter()
        outs = [ref.cuda_voxelization(pts[k], cfg["voxel_size"], cfg["point_cloud_range"], P, V) for k in range(frames)]
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / frames
        for k, (v, c, n) in enumerate(outs):
            m = m_list[k]
            same = same and v.size(0) == m and torch.equal(c, plan.coors[k, :m]) and torch.equal(n, plan.num_points[k, :m]) \
                and torch.equal(v, plan.voxels[k, :m])
        return {"ms_per_frame": round(dt * 1e3, 3), "value": round(cfg["n"] / dt / 1e6, 2), "unit": UNIT, "frames": frames,
                "same_output_as_ours": bool(same),
                "what": "reference hard_voxelize_gpu (voxelization_cuda.cu, unmodified, -arch sm_100a), host wall time per "
                        "frame incl. its own device syncs, same GPU"}
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:300]}


def run_latency_extras(cfg, pts, dev, reps=100):
    """Not the headline metric: what a detector that keeps the reference's call pattern sees.
    batch_B: device time per step of the pre-allocated batched call at B frames (DetMatch trains with
    B = 4 labeled + unlabeled frames per GPU, configs/detmatch/001/detmatch/split_0.py:108-112);
    forward_wall_us: host wall time of the unchanged per-frame API, Voxelization.forward (output
    allocation + launch sequence + the voxel_num read-back that sizes the returned views)."""
    import torch
    from detmatch_b200.ops import Voxelization
    from detmatch_b200.ops.voxel import HardVoxelizeBatchPlan
    N, C, P, V = cfg["n"], cfg["c"], cfg["max_num_points"], cfg["max_voxels"]
    out = {}
    for B in (1, 4, 8):
        if B > len(pts):
            continue
        plan = HardVoxelizeBatchPlan([N] * B, C, cfg["voxel_size"], cfg["point_cloud_range"], P, V, dev).bind(pts[:B])
        for _ in range(5):
            plan.run()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            plan.run()
        b.record()
        torch.cuda.synchronize(dev)
        out[f"batch_{B}_ms_per_step"] = round(a.elapsed_time(b) / reps, 4)
        del plan
    layer = Voxelization(cfg["voxel_size"], cfg["point_cloud_range"], P, V).eval()
    for _ in range(5):
        layer(pts[0])
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(reps):
        layer(pts[0])
    torch.cuda.synchronize(dev)
    out["forward_wall_us"] = round((time.perf_counter() - t0) / reps * 1e6, 1)
    return out


def run_fused_extras(cfg, pts, plan, dev, reps=50):
    """Not the headline metric: the same batch through the two SURVEY 8(f)-1 output modes, device
    time per step (CUDA events, rank 0).  packed = concatenated voxels / num_points / (b, z, y, x)
    coordinates written directly (what the detectors' voxelize() returns after torch.cat);
    packed_mean = the same with HardSimpleVFE folded into the expansion ((sum M, C) means instead of
    the (sum M, P, C) tensor)."""
    import ctypes

    import torch
    from detmatch_b200 import _cabi
    from detmatch_b200._torch_glue import ptr, stream_ptr
    F, N, C = len(pts), cfg["n"], cfg["c"]
    P, V = cfg["max_num_points"], cfg["max_voxels"]
    L = _cabi.lib()
    cap = F * min(N, V)
    coors = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    num = torch.empty((cap,), dtype=torch.int32, device=dev)
    vnum = torch.empty((F,), dtype=torch.int32, device=dev)
    pp = (ctypes.c_void_p * F)(*[t.data_ptr() for t in pts])
    nn = (ctypes.c_int64 * F)(*[N] * F)
    out = {}
    for name, mean in (("packed", 0), ("packed_mean", 1)):
        vox = torch.empty((cap, C) if mean else (cap, P, C), dtype=torch.float32, device=dev)

        def run():
            _cabi.check(L.pcfe_hard_voxelize_packed_batch_f32(pp, nn, F, C, plan.vs, plan.rg, None, P, V, mean, ptr(vox),
                                                              ptr(coors), ptr(num), cap, ptr(vnum), ptr(plan.ws),
                                                              plan.ws.numel(), dev.index, stream_ptr(dev)), name)
        for _ in range(5):
            run()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            run()
        b.record()
        torch.cuda.synchronize(dev)
        out[name + "_ms_per_step"] = round(a.elapsed_time(b) / reps, 4)
        del vox
    out["rows"] = int(vnum.sum().item())
    out["note"] = ("device time of pcfe_hard_voxelize_packed_batch_f32 on the same batch (mean = 0 / 1); "
                   "parity: tests/test_gpu_packed.py, tests/test_gpu_vfe.py")
    return out


def run_e2e(args, cfg, host, dev, total_frames, barrier):
    """Same metric through the package's public host-buffer call (detmatch_b200.ops.HostVoxelizePipeline,
    the pre-allocated form of voxelize_batch_host): every step copies this rank's frames from pinned host
    memory to the device, voxelizes them with the packed C-ABI call and brings the returned rows --
    voxels[:M], num_points[:M], (batch, z, y, x) coordinates, concatenated over the frames exactly as the
    reference's voxelize() loop returns them (voxelnet.py:60-67) -- into pinned host memory.  Wall
    clock around the calls, MAX over ranks."""
    import torch
    from detmatch_b200.ops import HostVoxelizePipeline
    N, C = cfg["n"], cfg["c"]
    pipe = HostVoxelizePipeline([N] * len(host), C, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"],
                                cfg["max_voxels"], device=dev, chunk=args.e2e_chunk)
    torch.cuda.synchronize(dev)
    pipe.run(host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        v, n, c = pipe.run(host)
    torch.cuda.synchronize(dev)
    dt = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
    assert v.size(0) == sum(pipe.counts) and n.size(0) == v.size(0) and c.size(0) == v.size(0)
    return {"value": round(total_frames * N / dt / 1e6, 1), "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
            "d2h_bytes_per_step": int(pipe.d2h_bytes), "ms_per_step": round(dt * 1e3, 3), "steps": args.e2e_steps,
            "rows": int(v.size(0)),
            "api": "detmatch_b200.ops.HostVoxelizePipeline.run (voxelize_batch_host): pinned host frames in, "
                   f"concatenated (voxels, num_points, coors_batch) out in pinned host memory; {args.e2e_chunk}-frame chunks "
                   "pipelined on H2D / compute / D2H streams, pcfe_hard_voxelize_packed_batch_f32 per chunk; bytes are "
                   "this rank's per step"}


def cpu_baseline_subprocess(args):
    """The reference arm on a bounded sample, in a fresh process (fork-safe: no CUDA there)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1",
           "--workload", args.workload, "--gpus", "1", "--sample-frames", "32"]
    try:
        env = dict(os.environ)
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
            env.pop(k, None)
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, env=env)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)["cpu_baseline"]
        return {"error": (out.stderr or "no output")[-300:]}
    except Exception as e:
        return {"error": str(e)[:300]}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args, sample_frames=args.sample_frames)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
