"""Summarise an .ncu-rep: key raw metrics per kernel launch + executed-instruction mix by opcode and by source line.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--lines 40]
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg"]
want += [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
want += [h for h in hdr if h.startswith("sm__inst_executed_pipe_") and h.endswith(".sum")]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in want:
        if k in d and d[k] not in ("", "n/a"):
            print(f"{k}: {d[k]} {units[hdr.index(k)]}")
    print("---")
    break
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
cur = None
curline = None
first_kernel_done = False
ops, lines, linesrc, tot = collections.Counter(), collections.Counter(), {}, 0
samples, stot = collections.Counter(), 0
n_func = 0
for r in rows:
    if len(r) == 2 and r[0] == "Function Name":
        continue
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        iI = hdr.index("Instructions Executed")
        iS = hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= iI:
        continue
    if r[0] != "":
        curline = int(r[0])
        linesrc[(cur, curline)] = r[1].strip()
        continue
    if r[iI] in ("-", ""):
        continue
    t = r[3].strip().split()
    op = t[1] if t[0].startswith("@") else t[0]
    n = int(r[iI])
    ops[op.split(".")[0]] += n
    lines[(cur, curline)] += n
    tot += n
    if r[iS] not in ("-", ""):
        samples[(cur, curline)] += int(r[iS])
        stot += int(r[iS])
print(f"# executed warp-instructions listed on the source page: {tot} (each SASS instruction appears once per listed launch)")
for op, n in ops.most_common(32):
    print(f"{op:12s} {100 * n / tot:5.1f}%")
print()
for (f, l), n in lines.most_common(nlines):
    print(f"{f}:{l:4d} {100 * n / tot:5.1f}% | {linesrc.get((f, l), '')[:120]}")

# optional: --ranges file:lo-hi:name,... sums executed instructions over source line ranges
if "--ranges" in sys.argv:
    spec = sys.argv[sys.argv.index("--ranges") + 1]
    px = float(sys.argv[sys.argv.index("--px") + 1]) if "--px" in sys.argv else None
    print()
    for item in spec.split(","):
        f, rng, name = item.split(":")
        lo, hi = (int(v) for v in rng.split("-"))
        n = sum(v for (ff, l), v in lines.items() if ff == f and lo <= l <= hi)
        sm = sum(v for (ff, l), v in samples.items() if ff == f and lo <= l <= hi)
        extra = f"  {32 * n / px:6.1f} lane-instr/px" if px else ""
        print(f"{name:28s} instr {100 * n / tot:5.1f}%{extra}   stall samples {100 * sm / max(stot, 1):5.1f}%")
