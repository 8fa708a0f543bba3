"""Instruction mix of one kernel from an .ncu-rep source page: python ncu_sass_mix.py rep regex"""
import csv, subprocess, sys, collections
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2]],
                     stdout=subprocess.PIPE, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr = rows[1]
ie, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
mix, smp = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= ie: continue
    try: n = float(r[ie]); s = float(r[ismp])
    except ValueError: continue
    toks = r[isrc].split()
    op = toks[0] if not toks[0].startswith("@") else toks[1]
    op = op.split(".")[0]
    mix[op] += n; smp[op] += s; tot += n
print("total warp instr", tot)
ts = sum(smp.values()) or 1
for op, n in mix.most_common(22):
    print(f"{op:10s} {n/tot:6.1%} instr   {smp[op]/ts:6.1%} samples")
