"""Per-thread dynamic instruction counts of one kernel split at its barriers, plus opcode mix.
Usage: python tests/native/ncu_regions.py report.ncu-rep kernel_regex warps_total"""
import csv, subprocess, sys, re

def main(path, kre, unit):
    out = subprocess.run(["ncu","-i",path,"--page","source","--csv","--kernel-name","regex:"+kre],stdout=subprocess.PIPE,text=True).stdout
    rows=list(csv.reader(out.splitlines()))
    hdr=None;data=[]
    for r in rows:
        if r and r[0]=="Address":
            if hdr is not None: break
            hdr=r
        elif hdr is not None and len(r)==len(hdr): data.append(r)
    iS,iI,iW,iT=(hdr.index(k) for k in ("Source","Instructions Executed","Warp Stall Sampling (All Samples)","Avg. Threads Executed"))
    tot=sum(int(r[iI]) for r in data); tots=sum(int(r[iW]) for r in data) or 1
    print("SASS lines",len(data),"warp-inst per warp",round(tot/unit,1))
    cur=0;cs=0;start=0;ops={};thr=0
    for k,r in enumerate(data):
        n=int(r[iI])/unit; cur+=n; cs+=int(r[iW]); thr+=n*float(r[iT])
        src=r[iS].strip(); t=src.split()
        op=(t[1] if src.startswith('@') else t[0]).split('.')[0]
        ops[op]=ops.get(op,0)+n
        if 'BAR.SYNC' in src or k==len(data)-1:
            print(f"  lines {start:4d}-{k:4d}: {cur:7.1f} warp-inst/warp  avg active lanes {thr/max(cur,1e-9):4.1f}  stall share {cs/tots*100:4.1f}%")
            cur=0;cs=0;start=k+1;thr=0
    print("  ops:", ", ".join(f"{o} {v:.0f}" for o,v in sorted(ops.items(), key=lambda x:-x[1])[:18]))

if __name__=="__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]))
