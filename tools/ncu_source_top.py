#!/usr/bin/env python3
"""Per-instruction stall sampling of one kernel of an ncu report (--import-source on), summarised:
  python tools/ncu_source_top.py report.ncu-rep KERNEL_ID [TOP] [--ctx N]
prints the sample total by stall reason, then the TOP SASS instructions by samples (with their dominant reasons,
executed count and N lines of context), then the samples aggregated over windows of 64 instructions."""
import csv
import io
import subprocess
import sys

rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else 25
ctx = int(sys.argv[sys.argv.index("--ctx") + 1]) if "--ctx" in sys.argv else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", ":::" + kid], capture_output=True, text=True).stdout
lines = out.splitlines()
print(lines[0][:160])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
h = rows[0]
body = []
for r in rows[1:]:          # the page may repeat the header for a second view: keep the first (SASS) section
    if r == h:
        break
    if len(r) == len(h):
        body.append(r)
rows = body
S = h.index("# Samples")
EX = h.index("Instructions Executed")
reasons = [i for i, n in enumerate(h) if n.startswith("stall_") and "(Not Issued)" not in n]
tot = sum(int(r[S]) for r in rows)
print("instructions:", len(rows), "samples:", tot, "warp-instructions executed:", sum(int(r[EX]) for r in rows))
agg = {h[i]: sum(int(r[i]) for r in rows) for i in reasons}
print("by reason:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 200 > tot))


def why(r):
    d = sorted(((int(r[i]), h[i][6:]) for i in reasons), reverse=True)[:3]
    return " ".join(f"{n}:{c}" for c, n in d if c)


order = sorted(range(len(rows)), key=lambda i: -int(rows[i][S]))[:top]
print(f"\ntop {top} instructions by samples")
for i in sorted(order):
    r = rows[i]
    if ctx:
        for j in range(max(0, i - ctx), i):
            print(f"        {j:5d}            {rows[j][1].strip()[:90]}")
    print(f"  {100 * int(r[S]) / tot:5.1f}% {i:5d} ex {int(r[EX]):7d}  {r[1].strip()[:70]:70s} {why(r)}")
print("\nsamples per window of 64 instructions")
for w in range(0, len(rows), 64):
    s = sum(int(r[S]) for r in rows[w:w + 64])
    ex = max(int(r[EX]) for r in rows[w:w + 64])
    if s * 100 > tot:
        a = {h[i]: sum(int(r[i]) for r in rows[w:w + 64]) for i in reasons}
        d = sorted(a.items(), key=lambda kv: -kv[1])[:3]
        print(f"  [{w:5d},{w + 64:5d}) {100 * s / tot:5.1f}%  max-exec {ex:7d}  " + " ".join(f"{k[6:]}:{v}" for k, v in d))

# barrier-separated phases of a CTA: every warp spends the same wall time in a phase (working or waiting at its closing
# barrier), so the share of samples between two BAR.SYNC instructions is that phase's share of the CTA's lifetime
if "--phases" in sys.argv:
    print("\nphases (instructions between consecutive BAR.SYNC; share of samples = share of CTA lifetime)")
    start = 0
    for i, r in enumerate(rows + [None]):
        if r is None or "BAR.SYNC" in r[1]:
            end = i + 1 if r is not None else i
            seg = rows[start:end]
            sm = sum(int(x[S]) for x in seg)
            if seg:
                a = {h[k]: sum(int(x[k]) for x in seg) for k in reasons}
                d = sorted(a.items(), key=lambda kv: -kv[1])[:4]
                exs = sorted(int(x[EX]) for x in seg)
                print(f"  [{start:5d},{end:5d}) {100 * sm / tot:5.1f}%  instr-exec sum {sum(exs):9d}  " + " ".join(f"{k[6:]}:{v}" for k, v in d if v))
            start = end
