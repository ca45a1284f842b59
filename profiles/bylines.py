"""Per-source-line executed-instruction table of an ncu report, binned by line ranges.
    python profiles/bylines.py rep.ncu-rep file.cu kernel_substr  "name:lo-hi" ...
"""
import collections, csv, io, re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
rep, cu, kname = sys.argv[1:4]
bins = [(b.split(":")[0], *map(int, b.split(":")[1].split("-"))) for b in sys.argv[4:]]
run = lambda cmd: subprocess.run(cmd, capture_output=True, text=True).stdout
src = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv"]))))
h = src[1]
ci, si, ss = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
data = [r for r in src[2:] if len(r) == len(h)]
cub = "/tmp/_bylines.cubin"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
                       f"-I{ROOT / 'include'}", f"-I{Path(cu).parent}", "-cubin", "-o", cub, cu])
fn, line, fname, ins = None, None, None, collections.defaultdict(list)
for l in run(["nvdisasm", "-g", "-c", cub]).splitlines():
    m = re.match(r"^\.text\.(\S+):", l)
    if m: fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: fname, line = Path(m.group(1)).name, int(m.group(2)); continue
    m = re.match(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and fn: ins[fn].append((fname, line, m.group(2).strip()))
cands = [k for k in ins if kname in k and len(ins[k]) == len(data)]
L = ins[cands[0]]
tot = sum(int(r[ci]) for r in data); ts = max(1, sum(int(r[ss]) for r in data))
nwarps = None
agg = collections.defaultdict(lambda: [0, 0, 0])
main = Path(cu).name
for a, (f, ln, txt) in zip(data, L):
    key = "other-file:" + f
    if f == main:
        key = "unbinned"
        for name, lo, hi in bins:
            if lo <= ln <= hi: key = name; break
    agg[key][0] += int(a[ci]); agg[key][1] += int(a[ss]); agg[key][2] += 1
print(f"total warp instr {tot}, samples {ts}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:28s} {100*v[0]/tot:5.1f}% instr  {100*v[1]/ts:5.1f}% samples  {v[2]:5d} SASS")
