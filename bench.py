#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric: M points/s of Waymo-shape hard voxelization
(config C4: 64 frames x 180 000 points x 5 features per GPU, voxel [0.1,0.1,0.15], range
[-75.2,-75.2,-2,75.2,75.2,4], max_points 5, max_voxels 150 000) on N B200s, next to the HBM
roofline and the reference's CPU op timed on the same host.

    python bench.py --gpus 1 --steps K --warmup W                 (our arm, one GPU)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N GPUs)
    python bench.py --impl reference --gpus N --steps K --warmup W    (reference CPU arm)

A step = one pass of the hot path over one batch of 64 synthetic frames per GPU.  Frames shard
across GPUs with no collective (weak scaling: every rank voxelizes its own 64 frames); the only
communication is a barrier and a MAX over the ranks' device times.  Prints ONE JSON line.
"""
import argparse
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
# DRAM traffic per 64-frame step from the committed ncu capture (profiles/r01_v14_ncu_summary.txt):
# hvb_bin 288.7 MB + hvb_bucket_rec 174.3 MB + hvb_scan_firsts 3.1 MB + hvb_expand_rec 1042.4 MB
NCU_TRAFFIC_BYTES_PER_STEP = {"C4": 1_510_900_000}
NCU_EXPAND_TRAFFIC = {"C4": 1_045_500_000}
NCU_PROFILE = "profiles/r01_v14_ncu_summary.txt"
KEPT_POINTS_PER_FRAME = {"C4": 156_000}  # points surviving the max_points cap (oracle, seed 4000)
METRIC = "hard_voxelize_throughput"
UNIT = "Mpoints/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU (default: the config's 64)")
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
    return ap.parse_args()


def workload_config(args):
    from detmatch_b200 import synth
    cfg = dict(synth.CONFIGS[args.workload])
    if args.frames > 0:
        cfg["frames"] = args.frames
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
def _ref_worker(conn, cfg, frame_id, use_ref):
    """One single-threaded worker: owns one frame, voxelizes it with the reference op on demand."""
    import torch
    torch.set_num_threads(1)
    from detmatch_b200 import synth
    if use_ref:
        from oracle import ref
        ref.module()
    else:
        from oracle import oracle
        oracle.lib()
    pts = synth.lidar_frame(cfg["n"], cfg["c"], synth.seed_for(cfg["index"], frame_id), cfg["r_max"])
    conn.send("ready")
    while True:
        cmd = conn.recv()
        if cmd != "run":
            break
        if use_ref:
            # voxelize.py:46-58 call pattern, including the three new_zeros
            v, c, n = ref.voxelization(pts, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], cfg["max_voxels"])
        else:
            v, c, n = oracle.hard_voxelize(pts.numpy(), cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], cfg["max_voxels"])
        conn.send(int(n.shape[0]))


def run_reference(args, quiet=False):
    """Times the reference's CPU hard_voxelize on this host: one single-threaded worker process
    per core, one frame per worker per step (a bounded sample of the workload)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    import multiprocessing as mp
    cfg = workload_config(args)
    from oracle import ref
    use_ref = ref.available()
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    procs = args.ref_procs if args.ref_procs > 0 else min(cores, 64)
    procs = max(1, min(procs, cfg["frames"] * max(args.gpus, 1)))
    ctx = mp.get_context("fork")
    workers = []
    for k in range(procs):
        parent, child = ctx.Pipe()
        p = ctx.Process(target=_ref_worker, args=(child, cfg, k, use_ref), daemon=True)
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
    pts_per_step = procs * cfg["n"]
    value = pts_per_step / t_step / 1e6
    kind = "reference" if use_ref else "port"
    sample = (f"{procs} frames/step ({cfg['n']} pts x {cfg['c']}), one single-threaded worker process per core, "
              f"{len(times)} steps; {'oracle/_ref = reference voxelization_cpu.cpp compiled in place' if use_ref else 'oracle C port'}")
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": max(args.warmup, 1), "ms_per_step": round(t_step * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": _config_dict(cfg, args, frames_per_step=procs),
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": procs, "kind": kind, "sample": sample,
                         "host_cores_visible": cores, "mean_voxels_per_frame": round(sum(ms) / len(ms), 1)},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not quiet:
        print(json.dumps(line), flush=True)
    return line


def _config_dict(cfg, args, frames_per_step):
    return {"workload": f"{args.workload}: Waymo-shape hard voxelization" if args.workload == "C4" else args.workload,
            "frames_per_gpu_per_step": frames_per_step, "points_per_frame": cfg["n"], "features": cfg["c"],
            "voxel_size": cfg["voxel_size"], "point_cloud_range": cfg["point_cloud_range"],
            "max_num_points": cfg["max_num_points"], "max_voxels": cfg["max_voxels"],
            "generator": "LiDAR-like (SURVEY 8(d)), seed = 1000*config + frame",
            "l2": "inputs+outputs per step (~0.94 GB) exceed the 126 MB L2; no explicit flush",
            "sharding": "frames, no collective"}


# --------------------------------------------------------------------------------------------
# multi-GPU plumbing (no data-path collective: frames shard, only times are reduced)
# --------------------------------------------------------------------------------------------
def frame_seeds(cfg_index, rank, frames_per_rank):
    """Global frame ids owned by `rank`: a contiguous block, disjoint across ranks."""
    return [rank * frames_per_rank + k for k in range(frames_per_rank)]


def max_over_ranks(value, device=None):
    """MAX of a python float over all ranks (identity when not distributed).  Works with nccl
    (device tensor) and gloo (CPU tensor)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
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
        dist.init_process_group("nccl", device_id=dev)
    cfg = workload_config(args)
    F, N, C = cfg["frames"], cfg["n"], cfg["c"]
    P, V = cfg["max_num_points"], cfg["max_voxels"]
    if args.hv_wave:
        _cabi.debug_set("hv_wave", args.hv_wave)
    if args.hv_bucket_avg:
        _cabi.debug_set("hv_bucket_avg", args.hv_bucket_avg)
    for kv in args.debug:
        name, _, val = kv.partition("=")
        _cabi.debug_set(name, int(val))

    # synthetic frames, generated on the host; each rank has its own 64 frames
    host = [synth.lidar_frame(N, C, synth.seed_for(cfg["index"], fid), cfg["r_max"]).pin_memory()
            for fid in frame_seeds(cfg["index"], rank, F)]
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
    ms_step = max_over_ranks(ms_total, dev) / args.steps
    total_points = world * F * N
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

    # ---- end to end through the public API with HOST buffers --------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, cfg, plan, host, pts, dev, world, barrier)

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
    algo = algorithmic_bytes(N, C, P, m_list)  # one rank's step
    achieved = algo / (ms_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": NCU_TRAFFIC_BYTES_PER_STEP.get(args.workload),
                "kernel": "hard-voxelize launch sequence per GPU (" + ", ".join(k for k in (kernels or {})) + "); "
                          "achieved = algorithmic bytes of the step / CUDA-event step time",
                "algorithmic_bytes_per_step": algo, "peak_source": peak_src + " (of measured)",
                "traffic_source": "sum of dram__bytes_read+write over the step's kernels, ncu --set full, " + NCU_PROFILE,
                "mean_voxels_per_frame": round(sum(m_list) / len(m_list), 1)}
    if kernels and "hvb_expand" in kernels:
        # dominant kernel (the expansion): writes every returned element once, reads the kept rows and
        # one first-point index per voxel (the records of multi-point voxels are not counted)
        kept = KEPT_POINTS_PER_FRAME.get(args.workload)
        m_tot = sum(m_list)
        k_bytes = m_tot * (P * C * 4 + 16) + m_tot * 4 + (kept * F * C * 4 if kept else 0)
        k_ms = kernels["hvb_expand"]["ms_per_step"]
        roofline["dominant_kernel"] = {"name": "hvb_expand", "ms_per_launch": k_ms,
                                       "algorithmic_bytes_per_launch": k_bytes,
                                       "achieved": round(k_bytes / (k_ms * 1e-3) / 1e9, 1),
                                       "frac": round(k_bytes / (k_ms * 1e-3) / 1e9 / peak, 4),
                                       "traffic": NCU_EXPAND_TRAFFIC.get(args.workload)}

    fused = None
    if not args.no_extras and P == 5 and C in (4, 5):
        fused = run_fused_extras(cfg, pts, plan, dev)

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        cpu_baseline = cpu_baseline_subprocess(args)

    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "ours",
        "config": _config_dict(cfg, args, frames_per_step=F),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks, "kernels": kernels, "fused": fused,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    F, N, C = cfg["frames"], cfg["n"], cfg["c"]
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


def run_e2e(args, cfg, plan, host, pts, dev, world, barrier):
    """Same metric through the public batched call with HOST buffers: every step copies the 64
    frames from pinned host memory to the device, voxelizes, reads voxel_num back and brings the
    returned rows of every frame -- voxels[:M], coors[:M], num_points[:M], concatenated over the
    frames exactly as the reference's voxelize() loop returns them (voxelnet.py:60-67) -- into
    pinned host memory.

    The batch is cut into chunks of 8 frames on three streams (H2D, compute, D2H) so that the
    upload of chunk i+1, the kernels of chunk i and the download of chunk i-1 overlap; the
    device-to-host size of a chunk is only known once its voxel_num has reached the host, which
    is the one host synchronisation per chunk.  The rows of a chunk are concatenated on the device
    (torch.cat into a staging buffer) and leave with three copies per chunk instead of three per
    frame: the copy engine no longer idles between 192 small transfers."""
    import torch
    from detmatch_b200.ops.voxel import HardVoxelizeBatchPlan
    F, N, C = cfg["frames"], cfg["n"], cfg["c"]
    P, V = cfg["max_num_points"], cfg["max_voxels"]
    CH = args.e2e_chunk if F % args.e2e_chunk == 0 else F
    chunks = [list(range(i, i + CH)) for i in range(0, F, CH)]
    plans = [HardVoxelizeBatchPlan([N] * CH, C, cfg["voxel_size"], cfg["point_cloud_range"], P, V, dev).bind(
        [pts[k] for k in ch]) for ch in chunks]
    cap_rows = F * min(V, N)
    out_vox = torch.empty((cap_rows, P, C), dtype=torch.float32).pin_memory()
    out_coors = torch.empty((cap_rows, 3), dtype=torch.int32).pin_memory()
    out_num = torch.empty((cap_rows,), dtype=torch.int32).pin_memory()
    stage = [(torch.empty((CH * min(V, N), P, C), dtype=torch.float32, device=dev),
              torch.empty((CH * min(V, N), 3), dtype=torch.int32, device=dev),
              torch.empty((CH * min(V, N),), dtype=torch.int32, device=dev)) for _ in range(2)]
    cnt_host = [torch.empty((CH,), dtype=torch.int32).pin_memory() for _ in chunks]
    s_in, s_comp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    h2d = F * N * C * 4
    offsets = []

    def step():
        ev_c = []
        for i, ch in enumerate(chunks):
            with torch.cuda.stream(s_in):
                for k in ch:
                    pts[k].copy_(host[k], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
            with torch.cuda.stream(s_comp):
                s_comp.wait_event(ev_in)
                plans[i].run()
                cnt_host[i].copy_(plans[i].voxel_num, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_comp)
                ev_c.append(ev)
        d2h = 0
        row0 = 0
        offsets.clear()
        for i, ch in enumerate(chunks):
            ev_c[i].synchronize()  # the caller needs M to size what it reads back
            counts = cnt_host[i].tolist()
            tot = sum(counts)
            d2h += len(ch) * 4 + tot * (P * C * 4 + 16)
            sv, sc, sn = stage[i & 1]
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_c[i])
                torch.cat([plans[i].voxels[j, :m] for j, m in enumerate(counts)], dim=0, out=sv[:tot])
                torch.cat([plans[i].coors[j, :m] for j, m in enumerate(counts)], dim=0, out=sc[:tot])
                torch.cat([plans[i].num_points[j, :m] for j, m in enumerate(counts)], dim=0, out=sn[:tot])
                out_vox[row0:row0 + tot].copy_(sv[:tot], non_blocking=True)
                out_coors[row0:row0 + tot].copy_(sc[:tot], non_blocking=True)
                out_num[row0:row0 + tot].copy_(sn[:tot], non_blocking=True)
            offsets.append((row0, counts))
            row0 += tot
        s_out.synchronize()
        return d2h

    torch.cuda.synchronize(dev)
    step()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.e2e_steps):
        d2h = step()
    torch.cuda.synchronize(dev)
    dt = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps, dev)
    # the downloaded rows are the device results (spot check, outside the timed region)
    row0, counts = offsets[-1]
    last0 = row0 + sum(counts[:-1])
    m = counts[-1]
    assert torch.equal(out_vox[last0:last0 + m], plans[-1].voxels[CH - 1, :m].cpu())
    assert torch.equal(out_coors[last0:last0 + m], plans[-1].coors[CH - 1, :m].cpu())
    return {"value": round(world * F * N / dt / 1e6, 1), "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": int(d2h), "ms_per_step": round(dt * 1e3, 3), "steps": args.e2e_steps,
            "api": "HardVoxelizeBatchPlan.run (pcfe_hard_voxelize_batch_f32), pinned host in/out buffers, "
                   f"{CH}-frame chunks pipelined on H2D / compute / D2H streams, rows concatenated over frames "
                   "on the device (as the reference's voxelize() returns them) before the read-back"}


def cpu_baseline_subprocess(args):
    """The reference arm on a bounded sample, in a fresh process (fork-safe: no CUDA there)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1",
           "--workload", args.workload, "--gpus", "1"]
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
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
