#!/usr/bin/env python
"""profiles/r02_<kernel>_traffic.json from an ncu --set full capture of bench.py's dominant kernel:
    ncu --set full --clock-control none -k regex:strip16 -s <launches of the warm-up> -c <launches per step> -o prof python bench.py ...
    python tools/ncu_traffic.py prof.ncu-rep k_metgrid_strip16 <kernel_points>
Sums dram__bytes_read / dram__bytes_write over the captured launches (= one step).  bench.py only uses the file when the
kernel name and the number of output points match the run."""
import csv, json, os, subprocess, sys
rep, kernel, points = sys.argv[1], sys.argv[2], float(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H, U = rows[0], rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0}
us, n = 0.0, 0
for r in rows[2:]:
    if kernel not in r[H.index("Kernel Name")]:
        continue
    n += 1
    for k in tot:
        tot[k] += float(r[H.index(k)].replace(",", "")) * scale[U[H.index(k)]]
    t = float(r[H.index("gpu__time_duration.sum")].replace(",", ""))
    us += t * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(U[H.index("gpu__time_duration.sum")], 1.0)
doc = {"kernel": kernel, "kernel_points": points, "launches": n, "dram_bytes_read": tot["dram__bytes_read.sum"],
       "dram_bytes_write": tot["dram__bytes_write.sum"], "ncu_duration_us_sum": us, "source": os.path.basename(rep)}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_%s_traffic.json" % kernel)
json.dump(doc, open(path, "w"), indent=1)
print(json.dumps(doc))
