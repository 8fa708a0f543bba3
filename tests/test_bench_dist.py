"""CPU tests of bench.py's host-side multi-GPU logic (world size 2, gloo backend): frame
sharding is disjoint, the time reduction is a MAX over ranks, and under torchrun the reference
arm runs on rank 0 only and prints exactly one JSON line."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seeds = bench.frame_seeds(4, rank, 64)
    # each rank's "device time": rank 1 is slower; every rank must see the max
    got = bench.max_over_ranks(1.0 + rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, (seeds[0], seeds[-1], got))
    if rank == 0:
        out.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_sharding_and_max():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29613, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][:2] == (0, 63) and res[1][:2] == (64, 127)  # disjoint contiguous blocks
    assert res[0][2] == 2.0 and res[1][2] == 2.0              # MAX over ranks on every rank


def test_max_over_ranks_without_process_group():
    assert bench.max_over_ranks(3.5) == 3.5


def test_reference_arm_under_torchrun_prints_one_line():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29614", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "1", "--workload", "C1", "--ref-procs", "2"]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 2
    assert d["value"] > 0 and d["unit"] == "Mpoints/s"


def test_strong_scaling_frame_blocks():
    """contiguous blocks of F / G frames per GPU, disjoint, covering the batch."""
    for total, world in ((64, 1), (64, 2), (64, 8), (128, 8), (10, 4)):
        blocks = [bench.frame_block(total, r, world) for r in range(world)]
        flat = [f for b in blocks for f in b]
        assert flat == list(range(total))
        assert max(len(b) for b in blocks) - min(len(b) for b in blocks) <= 1


def test_traffic_lookup_rejects_stale_capture(tmp_path, monkeypatch):
    """roofline.traffic comes from a committed ncu capture and is null when that capture does not
    contain the kernels the run launched."""
    prof = tmp_path / "profiles"
    prof.mkdir()
    (prof / "r99_traffic_C4.json").write_text(json.dumps({"workload": "C4", "source": "x", "kernels": {
        "hvb_bin_kernel": {"dram_read": 10, "dram_write": 1}, "hvb_expand_rec_kernel": {"dram_read": 5, "dram_write": 7}}}))
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    tot, per, src = bench.ncu_traffic("C4", ["hvb_bin", "hvb_expand", "hv_slow_fallback"])
    assert tot == 23 and "r99_traffic_C4.json" in src
    tot, per, src = bench.ncu_traffic("C4", ["hvb_bin", "hvb_bucket", "hvb_expand"])
    assert tot is None and "STALE" in src
    assert bench.ncu_traffic("C5", ["hvb_bin"])[0] is None
