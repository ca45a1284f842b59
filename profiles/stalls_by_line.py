"""Stall samples of one reason (e.g. stall_long_sb) by CUDA source line.
    python profiles/stalls_by_line.py rep.ncu-rep file.cu kernel_substr stall_long_sb [stall_barrier ...]
The .cu must be the version that was profiled (the SASS is re-generated from it to get line info)."""
import collections, csv, io, re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
rep, cu, kname = sys.argv[1:4]
reasons = sys.argv[4:]
run = lambda cmd: subprocess.run(cmd, capture_output=True, text=True).stdout
src = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv"]))))
h = src[1]
data = [r for r in src[2:] if len(r) == len(h)]
cub = "/tmp/_sbl.cubin"
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
assert cands, [(k[-30:], len(v)) for k, v in ins.items()]
L = ins[cands[0]]
text = Path(cu).read_text().split("\n")
for reason in reasons:
    ci = h.index(reason)
    agg = collections.Counter()
    for a, (f, ln, txt) in zip(data, L):
        agg[(f, ln)] += int(a[ci] or 0)
    tot = sum(agg.values())
    print(f"## {reason}: {tot} samples")
    for (f, ln), v in agg.most_common(14):
        s = text[ln - 1].strip()[:90] if f == Path(cu).name and 0 < ln <= len(text) else ""
        print(f"  {f:14s} L{ln:4d} {100 * v / max(1, tot):5.1f}% | {s}")
