"""Dev helper: STATIC SASS instruction count of one kernel per function of flat_core.cuh (code-size budget: the
hot loop must stay inside the 32 KB L1.5 instruction cache = 2048 instructions).
usage: python scripts/sass_regions.py object.o kernel_substring [source.cuh]"""
import collections, os, re, subprocess, sys, tempfile
obj, kname = sys.argv[1:3]
srcname = sys.argv[3] if len(sys.argv) > 3 else "flat_core.cuh"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
src = open(os.path.join(os.path.dirname(os.path.abspath(obj)), srcname)).read().splitlines()
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r"\s*HVP_HD\s+[\w:<>&\*\s]*?(\w+)\(", l)
    if m: marks.append((i, m.group(1)))
def region(l):
    name = "?"
    for a, nm in marks:
        if a <= l: name = nm
    return name
cnt = collections.Counter(); inside = False; cur = ("?", 0); total = 0
for ln in sass:
    if ln.startswith(".text."):
        inside = kname in ln; continue
    if not inside: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        cnt[region(cur[1]) if cur[0] == srcname else cur[0]] += 1; total += 1
print("total SASS instructions:", total, f"({total * 16 / 1024:.1f} KB)")
for k, v in cnt.most_common(30):
    print(f"{k:18s} {v:6d}")
