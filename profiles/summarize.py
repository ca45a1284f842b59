"""Summarise an ncu report (read here, no GPU needed): key raw metrics of the captured kernel, executed
instructions and stall samples by opcode, and by CUDA source line (ncu's SASS page joined with
`nvdisasm -g` line info of the same .cu compiled with the same flags).

    python profiles/summarize.py gpurun_out/prof.ncu-rep oceantransportmatrixbuilder.jl_b200/csrc/fused_v2.cu k_fused_v2 > profiles/xyz.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
rep, cu, kname = sys.argv[1], sys.argv[2], sys.argv[3]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
hdr, units = raw[0], raw[1]
print(f"# ncu summary of {rep}\n")
for row in raw[2:]:
    d = dict(zip(hdr, row))
    print(f"## launch: {d.get('Kernel Name', '')[:90]}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
    for h, u, v in zip(hdr, units, row):
        if h in KEYS:
            print(f"  {h:85s} {v:>16s} {u}")
    print()

src = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv"]))))
h = src[1]
ci, si, ss = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
data = [r for r in src[2:] if len(r) == len(h)]


def opcode(txt):
    t = txt.split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


byop, samp = collections.Counter(), collections.Counter()
for r in data:
    byop[opcode(r[si])] += int(r[ci])
    samp[opcode(r[si])] += int(r[ss])
tot, ts = sum(byop.values()), max(1, sum(samp.values()))
print(f"## executed warp instructions by opcode (total {tot}, {len(data)} SASS instructions, {ts} samples)")
for op, n in byop.most_common(22):
    print(f"  {op:10s} {100 * n / tot:5.1f}% of instructions   {100 * samp[op] / ts:5.1f}% of stall samples")

cub = "/tmp/_summarize.cubin"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
                       f"-I{ROOT / 'include'}", f"-I{Path(cu).parent}", "-cubin", "-o", cub, cu])
fn, line, fname, ins = None, None, None, collections.defaultdict(list)
for l in run(["nvdisasm", "-g", "-c", cub]).splitlines():
    m = re.match(r"^\.text\.(\S+):", l)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        fname, line = Path(m.group(1)).name, int(m.group(2))
        continue
    m = re.match(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and fn:
        ins[fn].append((fname, line, m.group(2).strip()))
cands = [k for k in ins if kname in k and len(ins[k]) == len(data)]
if not cands:
    print(f"\n(no function matching {kname} with {len(data)} instructions; line join skipped: "
          f"{[(k[-40:], len(v)) for k, v in ins.items()]})")
    sys.exit(0)
L = ins[cands[0]]
agg = collections.defaultdict(lambda: [0, 0])
for a, (f, ln, txt) in zip(data, L):
    assert opcode(a[si]) == opcode(txt), "SASS mismatch: compile flags differ from the profiled build"
    agg[(f, ln)][0] += int(a[ci])
    agg[(f, ln)][1] += int(a[ss])
text = {}
for f in {k[0] for k in agg}:
    p = Path(cu).parent / f
    text[f] = p.read_text().split("\n") if p.exists() else []
print(f"\n## by CUDA source line ({cands[0][-60:]})")
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0] - 3000 * kv[1][1])[:50]:
    s = text.get(f, [])
    s = s[ln - 1].strip()[:100] if 0 < ln <= len(s) else ""
    print(f"  {f:12s} L{ln:4d} {100 * v[0] / tot:5.1f}% instr {100 * v[1] / ts:5.1f}% samples | {s}")
