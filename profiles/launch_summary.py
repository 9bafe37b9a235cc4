#!/usr/bin/env python3
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`): per-kernel mean duration
and share of the step.  usage: python profiles/launch_summary.py launches.csv [lanes]
(lanes = stream ranges per step: the DSP kernels and the output heads are launched `lanes` times per step; the network
kernels -- conv1, conv2, three GRU layers, or the fused k_net -- once per step over the whole batch since round 2's
"ranges"; pass a negative value for a round-1 capture, where every kernel ran once per lane)"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
H = rows[hdr]
ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
d = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else v * 1e3 if r[ui] == 'ms' else v
    d.setdefault(r[ki].split('(')[0], []).append(v)
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
NET = ('k_conv1', 'k_conv2', 'void k_tc2<0>', 'k_gru', 'k_gru_tc', 'void k_tc2<1>', 'k_net')
def mult(k):
    per_lane = abs(lanes) if (lanes < 0 or k not in NET) else 1
    return per_lane * (3 if k in ('k_gru', 'k_gru_tc', 'void k_tc2<1>') else 1)
per_step = {k: sum(v) / len(v) * mult(k) for k, v in d.items() if 'k_' in k}
tot = sum(per_step.values())
print(f"{'kernel':16s} {'launches':>8s} {'mean us':>10s} {'us/step':>10s} {'share':>7s}")
for k, v in d.items():
    if 'k_' in k:
        print(f"{k:16s} {len(v):8d} {sum(v)/len(v):10.1f} {per_step[k]:10.1f} {per_step[k]/tot:7.1%}")
print(f"{'sum':16s} {'':8s} {'':10s} {tot:10.1f}")
