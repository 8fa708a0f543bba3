"""Per-source-line instruction / stall shares of one kernel from an .ncu-rep captured with
--import-source on (run here, no GPU): python tests/native/ncu_lines.py REP KERNEL_REGEX [min_pct]"""
import csv, subprocess, sys, io

rep, kern = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                      "--kernel-name", "regex:" + kern, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Line No")
ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples")
ti = hdr.index("Thread Instructions Executed")
lines = [r for r in rows if r and r[0].isdigit() and len(r) > ii and r[ii].isdigit() and r[si].isdigit() and r[ti].isdigit()]
tot = sum(int(r[ii]) for r in lines); stot = sum(int(r[si]) for r in lines)
print("kernel", kern, "warp-instructions", tot, "samples", stot)
for r in lines:
    n = int(r[ii])
    if 100.0 * n / tot >= min_pct or 100.0 * int(r[si]) / max(stot, 1) >= min_pct:
        act = int(r[ti]) / max(n, 1)
        print("%5s inst %5.1f%% smp %5.1f%% act %4.1f | %s" % (r[0], 100.0 * n / tot, 100.0 * int(r[si]) / max(stot, 1), act, r[1].strip()[:120]))
