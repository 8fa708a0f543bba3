"""Instruction / stall-sample shares of one kernel per source file and per line range.
python tests/native/ncu_sections.py REP KERNEL_REGEX FILE_SUFFIX a:b:name [a:b:name ...]"""
import collections, csv, io, subprocess, sys

rep, kern, suffix = sys.argv[1:4]
secs = [(int(a), int(b), n) for a, b, n in (s.split(":") for s in sys.argv[4:])]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      "regex:" + kern, "--launch-count", "1"], capture_output=True, text=True).stdout
cur, data = None, []
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "File Path":
        cur = r[1]
    if r and r[0].isdigit() and len(r) > 8 and r[7].isdigit() and r[6].isdigit():
        data.append((cur, int(r[0]), int(r[7]), int(r[6])))
tot = sum(d[2] for d in data); st = sum(d[3] for d in data) or 1
byf, sf = collections.Counter(), collections.Counter()
for f, l, n, s in data:
    byf[f] += n; sf[f] += s
print("warp-instructions", tot, "samples", st)
for f in byf:
    print("  file %-60s inst %5.1f%% smp %5.1f%%" % (f[-60:], 100 * byf[f] / tot, 100 * sf[f] / st))
for a, b, name in secs:
    n = sum(d[2] for d in data if d[0].endswith(suffix) and a <= d[1] < b)
    s = sum(d[3] for d in data if d[0].endswith(suffix) and a <= d[1] < b)
    print("  %-16s inst %5.1f%% (%6.1f M) smp %5.1f%%" % (name, 100 * n / tot, n / 1e6, 100 * s / st))
