"""Turns the raw artefacts of tools/r2_profiles.sh (gpurun_out/r02_*) into the tracked summaries under profiles/.
    python tools/make_profiles_r02.py
"""
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

WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg"]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def summarise(rep, title):
    hdr, units, rows = raw_rows(rep)
    want = WANT + [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
    lines = [f"# {title}"]
    recs = []
    for r in rows:
        d = dict(zip(hdr, r))
        for k in want:
            if k in d and d[k] not in ("", "n/a"):
                lines.append(f"{k}: {d[k]} {units[hdr.index(k)]}")
        lines.append("---")
        recs.append((d, units, hdr))
    return lines, recs


def to_bytes(d, units, hdr, key):
    v, u = float(d[key].replace(",", "")), units[hdr.index(key)]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


# ---- front kernel: 9-frame launch (the batch pipeline's shape) and 4-frame launch (round 1's capture shape) ----
for tag, frames in (("9f", 9), ("4f", 4)):
    lines, recs = summarise(G / f"r02_front3_{tag}.ncu-rep",
                            f"ncu --set full --import-source on --clock-control none -k regex:front -c 1: front3_kernel<5,TMA,DIV=1,64>, one launch of {frames} frames 3840x2160")
    (P / f"r02_front3_{tag}_ncu_full.txt").write_text("\n".join(lines) + "\n")
    if tag == "9f":
        d, units, hdr = recs[0]
        traffic = int(to_bytes(d, units, hdr, "dram__bytes_read.sum") + to_bytes(d, units, hdr, "dram__bytes_write.sum"))
        px = frames * 3840 * 2160
        (P / "traffic.json").write_text(json.dumps({
            "front_kernel_dram_bytes_per_launch": traffic, "front_kernel_dram_bytes_per_px": round(traffic / px, 4),
            "source": f"profiles/r02_front3_9f_ncu_full.txt (front3_kernel: dram__bytes_read.sum + dram__bytes_write.sum, one launch = {frames} frames "
                      f"3840x2160 = {px / 1e6:.2f} Mpix; part of the launch's class map is still in the 126 MB L2 when the launch ends)",
            "algorithmic_bytes_per_launch": 2 * px}, indent=1) + "\n")
        print("front traffic", traffic, traffic / px)

# ---- the fused-BGR variant (9 interleaved B,G,R frames) ----
if (G / "r02_front3_bgr_9f.ncu-rep").exists():
    lines, recs = summarise(G / "r02_front3_bgr_9f.ncu-rep",
                            "ncu --set full --import-source on --clock-control none -k regex:front3_kernel -s 5 -c 1 python tools/bgr_probe.py --frames 9: "
                            "front3_kernel<5,TMA,DIV=1,64,BGR>, one launch of 9 interleaved B,G,R frames 3840x2160 (4 B/px algorithmic: 3 in + 1 out)")
    (P / "r02_front3_bgr_9f_ncu_full.txt").write_text("\n".join(lines) + "\n")

# ---- hysteresis kernels: in the pipeline (caches not flushed) vs cold ----
for tag, title in (("inpipe", "--cache-control none: what the kernels read was just written by the front kernel of the same chunk (27 frames = 3 chunks in flight)"),
                   ("cold", "default cache control (flushed before every replay): round 1's view")):
    lines, recs = summarise(G / f"r02_hyst_{tag}.ncu-rep", f"ncu --set full --clock-control none -k regex:ccl_sparse, {title}")
    extra = ["", "# per launch: DRAM bytes (read + write), L2 sector hit rate, duration"]
    for d, units, hdr in recs:
        hit = d.get("lts__t_sector_hit_rate.pct", "n/a")
        extra.append(f"{d['Kernel Name'][:60]:60s} dram {to_bytes(d, units, hdr, 'dram__bytes_read.sum') + to_bytes(d, units, hdr, 'dram__bytes_write.sum'):12.0f} B   "
                     f"l2 hit rate {hit} %   {d['gpu__time_duration.sum']} {units[hdr.index('gpu__time_duration.sum')]}")
    (P / f"r02_hysteresis_{tag}_ncu_full.txt").write_text("\n".join(lines + extra) + "\n")
    print("\n".join(extra))

# ---- launch list + shares ----
rows = [r for r in csv.reader(open(G / "r02_launches.csv")) if len(r) > 10]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    agg[r[ik].split("(")[0]].append(float(r[iv].replace(",", "")))
pipe = {k: v for k, v in agg.items() if not any(s in k for s in ("synth", "count255", "div1_check", "div3_check", "hash255"))}
tot = sum(sum(v) for v in pipe.values())
out = ["# per-kernel device time of `bench.py --frames 18 --steps 2 --warmup 3 --no-e2e --no-cpu --no-bands --no-extras` under ncu "
       "(--metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: shares, not absolutes)"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    share = f"{sum(v) / tot:.3f}" if k in pipe else "  -  "
    out.append(f"{k[:64]:64s} launches={len(v):3d} total_us={sum(v) / 1e3:9.1f} share_of_pipeline={share} avg_us={sum(v) / len(v) / 1e3:8.1f}")
(P / "r02_launch_shares.txt").write_text("\n".join(out) + "\n")
shutil.copy(G / "r02_launches.csv", P / "r02_launches.csv")
print("\n".join(out[:6]))

# ---- bench lines, config matrix ----
for name in ("r02_bench_n1.json", "r02_bench_reference_arm.json", "r02_config_matrix.json"):
    shutil.copy(G / name, P / name)
for n in (2, 8):
    src = G / f"r2c_bench_n{n}.json"
    if src.exists():
        shutil.copy(src, P / f"r02_bench_n{n}.json")
for n in (2, 8):
    for tag in ("check", "check5"):
        src = G / f"r2c_{tag}_n{n}.log"
        if src.exists():
            keep = [l for l in src.read_text().splitlines() if l.startswith("{")]
            (P / f"r02_multigpu_bands_{tag}_n{n}.jsonl").write_text("\n".join(keep) + "\n")

# ---- probes ----
for src, dst in (("r02_chunk_sweep.txt", "r02_chunk_sweep.txt"), ("r02_pytest_gpu.log", "r02_pytest_gpu.log")):
    if (G / src).exists():
        shutil.copy(G / src, P / dst)
probe = {}
for name in ("bgr_fused", "bgr_separate", "bgr8k_fused", "bgr8k_separate", "pdl0", "pdl1"):
    f = G / f"r02_{name}.json"
    if f.exists():
        try:
            probe[name] = json.loads(f.read_text().strip().splitlines()[-1])
        except Exception as e:
            probe[name] = {"error": str(e)}
if probe:
    (P / "r02_bgr_pdl_probes.json").write_text(json.dumps(probe, indent=1) + "\n")

# ---- SASS evidence ----
sass = subprocess.run([sys.executable, str(ROOT / "tools" / "sass_check.py"), "front3_kernelILi5ELb1ELi1ELi64ELb0"], capture_output=True, text=True).stdout
(P / "r02_front3_sass_summary.txt").write_text(sass)
lib = ROOT / "canny_edge_b200" / "libcanny_b200.so"
full = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN2cb2f313front3_kernelILi5ELb1ELi1ELi64ELb0EEEvNS_11FrontParamsE14CUtensorMap_st", str(lib)],
                      capture_output=True, text=True).stdout
keep = []
for line in full.splitlines():
    if "/*" in line and line.strip().startswith("/*") and line.count("/*") >= 2:
        keep.append(line.split("/*", 2)[0] + "/*" + line.split("/*", 2)[1].rstrip())   # address + instruction, encoding stripped
(P / "r02_front3_kernel.sass").write_text("\n".join(l.rstrip() for l in keep) + "\n")
print("sass lines", len(keep))
