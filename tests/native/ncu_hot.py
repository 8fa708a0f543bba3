"""Hot SASS lines of one kernel from an .ncu-rep captured with --import-source on.
Usage: python tests/native/ncu_hot.py report.ncu-rep kernel_regex [min_pct]"""
import csv
import subprocess
import sys


def main(path, kre, min_pct=0.6):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                         stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    data = []
    for r in rows:
        if r and r[0] == "Address":
            if hdr is not None:
                break  # first kernel instance only
            hdr = r
        elif hdr is not None and len(r) == len(hdr):
            data.append(r)
    iS, iI, iW, iT = (hdr.index(k) for k in ("Source", "Instructions Executed", "Warp Stall Sampling (All Samples)", "Avg. Threads Executed"))
    tot = sum(int(r[iI]) for r in data)
    tots = sum(int(r[iW]) for r in data) or 1
    print(f"total warp-instructions {tot}, stall samples {tots}, SASS lines {len(data)}")
    for k, r in enumerate(data):
        n, s = int(r[iI]), int(r[iW])
        if n > tot * min_pct / 100 or s > tots * 2 * min_pct / 100:
            print(f"{k:4d} inst {n / tot * 100:5.2f}%  stall {s / tots * 100:5.2f}%  thr {float(r[iT]):5.1f}  {r[iS].strip()[:84]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 0.6)
