"""Dev helper: per-SOURCE-LINE stall samples of one kernel from an ncu report.
ncu's CSV source page is per SASS instruction; nvdisasm -g gives the source line of every SASS instruction of the
same cubin.  The two listings are in the same order, so they are joined by instruction index.
usage: python scripts/ncu_lines.py report.ncu-rep object.o kernel_substring [top]"""
import collections
import csv
import io
import re
import subprocess
import sys
import tempfile
import os

rep, obj, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
lines, cur, inside = [], ("?", 0), False
for ln in sass:
    if ln.startswith(".text."):
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
if len(lines) != len(data):
    print(f"warning: {len(lines)} SASS instructions in the object vs {len(data)} in the report (different build?)")
def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0
keys = ["# Samples", "stall_long_sb", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_selected",
        "Instructions Executed", "Thread Instructions Executed"]
agg = collections.defaultdict(lambda: [0.0] * len(keys))
for i, r in enumerate(data[:len(lines)]):
    a = agg[lines[i]]
    for j, k in enumerate(keys):
        a[j] += f(r, k)
tot = sum(a[0] for a in agg.values())
print("file:line  samples%  long_sb wait short_sb branch selected | warp-instr  lanes/instr")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:<5d} {100 * a[0] / tot:5.1f}%  {int(a[1]):6d} {int(a[2]):6d} {int(a[3]):6d} {int(a[4]):6d} {int(a[5]):6d} | "
          f"{int(a[6]):10d}  {a[7] / max(a[6], 1):5.1f}")
