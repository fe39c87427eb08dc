#!/usr/bin/env python3
"""DRAM traffic of the dominant kernel from an `ncu --set full` report (runs here, no GPU):
    python tools/ncu_traffic.py gpurun_out/r02x_stages.ncu-rep 'k_fast_cells' 32 profiles/r02_fast_cells_traffic.json
sums dram__bytes_read.sum / dram__bytes_write.sum over every launch whose name matches the regex (the fused FAST kernel is
two launches per batch: class A and class B tiles) and records how many frames the launches covered.  bench.py reads the
JSON for `roofline.traffic` -- the number is never typed into bench.py."""
import csv, io, json, re, subprocess, sys

rep, pattern, frames, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ki, ri, wi, ti = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd = wr = us = 0.0
names = []
for r in rows[2:]:
    if re.search(pattern, r[ki]):
        rd += float(r[ri].replace(",", "")) * scale[units[ri]]
        wr += float(r[wi].replace(",", "")) * scale[units[wi]]
        us += float(r[ti].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[ti], 1.0)
        names.append(r[ki].split("(")[0])
assert names, f"no kernel matches {pattern}"
json.dump({"kernel_regex": pattern, "launches": names, "frames_per_launch": frames, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "duration_us_under_ncu": us, "source": f"ncu --set full --clock-control none, {rep.split('/')[-1]} (tools/ncu_traffic.py)"},
          open(out, "w"), indent=1)
print(open(out).read())
