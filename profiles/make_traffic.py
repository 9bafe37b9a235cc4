#!/usr/bin/env python3
"""profiles/traffic.json from an `ncu --page raw --csv` export of one `--set full` capture: per kernel,
dram__bytes_read.sum + dram__bytes_write.sum and the duration of one launch (mean over the captured
launches of that kernel).  usage: python profiles/make_traffic.py raw.csv <lanes> <source note> > profiles/traffic.json"""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
H, U = rows[0], rows[1]
ki, ri, wi, di = (H.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tscale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}
acc = collections.OrderedDict()
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "")
    key = {"k_tc2<1>": "k_gru", "k_tc2<0>": "k_conv2"}.get(name, "k_heads" if name.startswith("k_heads2") else name)
    b = float(r[ri].replace(",", "")) * scale[U[ri]] + float(r[wi].replace(",", "")) * scale[U[wi]]
    acc.setdefault(key, []).append((b, float(r[di].replace(",", "")) * tscale[U[di]]))
out = {"lanes": int(sys.argv[2])}
for k, v in acc.items():
    out[k] = {"dram_bytes_per_launch": sum(x[0] for x in v) / len(v), "duration_us": sum(x[1] for x in v) / len(v),
              "launches_captured": len(v), "source": sys.argv[3]}
print(json.dumps(out, indent=1))
