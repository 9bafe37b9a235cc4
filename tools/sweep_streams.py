#!/usr/bin/env python3
"""Stream-count sweep (BASELINE.json configs[4]): runs bench.py for a list of S on one GPU and prints
one table row per S: frames/s device-resident, end-to-end, ms/step, pipeline roofline fraction, and the
unmodified reference on the host cores for the same S.
usage: python tools/sweep_streams.py [--no-cpu] [S ...] > profiles/sweep.md"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
no_cpu = "--no-cpu" in sys.argv[1:]
extra = []
for i, x in enumerate(sys.argv[1:]):
    if x == "--model":
        extra = ["--model", sys.argv[i + 2]]
sizes = [int(x) for x in sys.argv[1:] if x.isdigit()] or [64, 256, 1024, 4096, 16384, 65536, 262144]
print("| streams S | device-resident frames/s | e2e frames/s | ms/step | pipeline GB/s (frac of HBM) | reference CPU frames/s (threads) | e2e speed-up |")
print("|---:|---:|---:|---:|---:|---:|---:|")
for S in sizes:
    steps = 300 if S <= 16384 else 100 if S <= 65536 else 40
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--streams", str(S), "--steps", str(steps), "--warmup", "20"]
                       + extra + (["--no-cpu-baseline"] if no_cpu else []), capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(f"| {S} | failed: {r.stderr[-200:]} |")
        continue
    d = json.loads(line[-1])
    cpu = d.get("cpu_baseline") or {}
    p = d["roofline"]["pipeline"]
    sp = d["e2e"]["value"] / cpu["value"] if cpu.get("value") else float("nan")
    print(f"| {S} (lanes {d.get('lanes', 1)}) | {d['value']:.3e} | {d['e2e']['value']:.3e} | {d['ms_per_step']:.4f} | {p['achieved_GBps']:.0f} ({p['frac']:.1%}) | "
          f"{cpu.get('value', float('nan')):.3e} ({cpu.get('cores', '?')}) | {sp:.1f}x |", flush=True)
