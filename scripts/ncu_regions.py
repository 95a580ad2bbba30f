"""Dev helper: warp-instructions, lanes per instruction and stall samples of one kernel from an ncu report, aggregated
by CODE REGION of flat_core.cuh (line ranges below) and by SASS opcode.
usage: python scripts/ncu_regions.py report.ncu-rep object.o kernel_substring solves_per_launch"""
import collections, csv, io, re, subprocess, sys, tempfile, os
rep, obj, kname = sys.argv[1:4]
solves = float(sys.argv[4]) if len(sys.argv) > 4 else 655360.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
lines, cur, inside = [], ("?", 0), False
ops = []
for ln in sass:
    if ln.startswith(".text."):
        inside = kname in ln; continue
    if not inside: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(\S+)", ln)
    if m:
        lines.append(cur); ops.append(m.group(1))
print("SASS instructions:", len(lines), "report rows:", len(data))
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
# function -> line range in flat_core.cuh: regenerate with  grep -n "HVP_HD" flat_core.cuh
src = open(os.path.join(os.path.dirname(os.path.abspath(obj)), "flat_core.cuh")).read().splitlines()
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r"\s*HVP_HD\s+[\w:<>&\*\s]*?(\w+)\(", l)
    if m: marks.append((i, m.group(1)))
def region(l):
    name = "?"
    for a, nm in marks:
        if a <= l: name = nm
    return name
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
opagg = collections.defaultdict(lambda: [0.0, 0.0])
for i, r in enumerate(data[:len(lines)]):
    fl, l = lines[i]
    name = region(l) if fl == "flat_core.cuh" else fl
    a = agg[name]; a[0] += f(r, "# Samples"); a[1] += f(r, "Instructions Executed"); a[2] += f(r, "Thread Instructions Executed")
    op = ops[i].split('.')[0].rstrip(';')
    opagg[op][0] += f(r, "Instructions Executed"); opagg[op][1] += f(r, "Thread Instructions Executed")
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:16s} samples {100*a[0]/ts:5.1f}%  warp-instr {100*a[1]/ti:5.1f}% ({a[1]/solves:8.0f}/solve) lanes {a[2]/max(a[1],1):5.1f}")
print("total warp-instr/solve", ti / solves, "thread-instr/solve", sum(a[2] for a in agg.values()) / solves)
for k, a in sorted(opagg.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{k:10s} {100*a[0]/ti:5.1f}%  lanes {a[1]/max(a[0],1):5.1f}")
