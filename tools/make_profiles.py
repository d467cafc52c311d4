"""Turns the raw artefacts of a GPU run (gpurun_out/) into the tracked summaries under profiles/.
    python tools/make_profiles.py r01_final
expects gpurun_out/{prof_final.ncu-rep, launches2.csv, bench.json, bench_ref.json}."""
import collections
import csv
import io
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
G, P = ROOT / "gpurun_out", ROOT / "profiles"
tag = sys.argv[1] if len(sys.argv) > 1 else "r01_final"

raw = subprocess.run(["ncu", "-i", str(G / "prof_final.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max"]
want += [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
lines, traffic = [], None
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in want:
        if k in d and d[k] not in ("", "n/a"):
            lines.append(f"{k}: {d[k]} {units[hdr.index(k)]}")
    lines.append("---")
    if "front2_kernel" in d["Kernel Name"] and traffic is None:
        def to_bytes(key):
            v, u = float(d[key].replace(",", "")), units[hdr.index(key)]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        traffic = int(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))
(P / f"{tag}_ncu_full.txt").write_text("\n".join(lines) + "\n")
(P / "traffic.json").write_text(json.dumps({
    "front_kernel_dram_bytes_per_launch": traffic,
    "front_kernel_dram_bytes_per_px": round(traffic / 33177600.0, 4),
    "source": f"profiles/{tag}_ncu_full.txt (front2_kernel: dram__bytes_read.sum + dram__bytes_write.sum, one launch = 4 frames 3840x2160 = "
              "33.18 Mpix; part of a launch's 33 MB class map is still in the 126 MB L2 when the launch ends)",
    "algorithmic_bytes_per_launch": 66355200}, indent=1) + "\n")

rows = [r for r in csv.reader(open(G / "launches2.csv")) if len(r) > 10]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    agg[r[ik].split("(")[0]].append(float(r[iv].replace(",", "")))
pipe = {k: v for k, v in agg.items() if not any(s in k for s in ("synth", "count255", "div1_check", "div3_check"))}
tot = sum(sum(v) for v in pipe.values())
out = ["# per-kernel device time of `bench.py --frames 18 --steps 2 --warmup 3 --no-e2e --no-cpu` under ncu (cold-cache, serialised: shares, not absolutes)"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    share = f"{sum(v) / tot:.3f}" if k in pipe else "  -  "
    out.append(f"{k[:64]:64s} launches={len(v):3d} total_us={sum(v) / 1e3:9.1f} share_of_pipeline={share} avg_us={sum(v) / len(v) / 1e3:8.1f}")
(P / f"{tag}_launch_shares.txt").write_text("\n".join(out) + "\n")
shutil.copy(G / "launches2.csv", P / f"{tag}_launches.csv")
shutil.copy(G / "bench.json", P / "r01_bench_n1.json")
if (G / "bench_ref.json").exists():
    shutil.copy(G / "bench_ref.json", P / "r01_bench_reference_arm.json")
mix = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(G / "prof_final.ncu-rep"), "--lines", "0"], capture_output=True, text=True).stdout
(P / f"{tag}_front2_opcode_mix.txt").write_text(mix[mix.index("# executed warp-instructions"):])
sass = subprocess.run([sys.executable, str(ROOT / "tools" / "sass_check.py"), "front2_kernelILi5ELb1ELi1ELi64"], capture_output=True, text=True).stdout
(P / "r01_front2_sass_summary.txt").write_text(sass)
print("\n".join(out[:6]))
print("traffic", traffic)
