"""DRAM traffic per kernel of one step from an `ncu --set full` report -> profiles/<tag>_traffic_<workload>.json,
the file bench.py's roofline.traffic is read from.  Launches of the same kernel are averaged.
    python tests/native/ncu_traffic.py REP WORKLOAD TAG "<the command the report was captured with>" FRAMES_PER_STEP """
import collections
import csv
import json
import os
import re
import subprocess
import sys

rep, workload, tag, cmd = sys.argv[1:5]
frames = int(sys.argv[5]) if len(sys.argv) > 5 else 64
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ix = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
for r in rows[2:]:
    m = re.search(r"(\w+_kernel)", r[ix["Kernel Name"]])
    name = m.group(1) if m else r[ix["Kernel Name"]][:40]
    a = acc[name]
    a[0] += float(r[ix["dram__bytes_read.sum"]]) * scale[units[ix["dram__bytes_read.sum"]]]
    a[1] += float(r[ix["dram__bytes_write.sum"]]) * scale[units[ix["dram__bytes_write.sum"]]]
    a[2] += float(r[ix["gpu__time_duration.sum"]])
    a[3] += 1
d = {"workload": workload, "frames_per_step": frames, "source": os.path.basename(rep) + " (" + cmd + ")",
     "kernels": {k: {"dram_read": int(v[0] / v[3]), "dram_write": int(v[1] / v[3]), "us": round(v[2] / v[3], 2), "launches": v[3]}
                 for k, v in acc.items()}}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "profiles",
                    f"{tag}_traffic_{workload}.json")
json.dump(d, open(path, "w"), indent=1)
print(path, json.dumps(d["kernels"]))
