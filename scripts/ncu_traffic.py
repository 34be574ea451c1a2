"""DRAM traffic per launch of every kernel in an .ncu-rep (--set full) -> profiles/r2_kernel_traffic.json, the file
bench.py reads its roofline.traffic from.   python scripts/ncu_traffic.py <rep> <workload>/<mode>"""
import csv
import json
import os
import re
import subprocess
import sys

rep, key = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ir, iw, iname = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for r in rows[2:]:
    short = re.sub(r"^(void )?(pph::)?(\(anonymous namespace\)::)?", "", r[iname]).split("<")[0].split("(")[0]
    byts = float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]]
    acc.setdefault(short, []).append(byts)
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_kernel_traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[key] = {k: (sum(v) / len(v) if len(v) == 1 else max(v)) for k, v in acc.items()}
data[key + "/launches"] = {k: v for k, v in acc.items()}
data["_source"] = "ncu --set full --clock-control none, one eager step (cold-ish caches), dram__bytes_read.sum + dram__bytes_write.sum per launch; kernels launched more than once per step: the largest launch"
json.dump(data, open(path, "w"), indent=1)
print(json.dumps(data[key], indent=1))
