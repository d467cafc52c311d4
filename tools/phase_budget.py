"""Per-phase instruction budget of a front kernel from an ncu capture: splits the SASS listing at BAR.SYNC instructions and
prints, per segment, executed warp-instructions (total and per pixel), stall samples and the opcode mix.
    python tools/phase_budget.py gpurun_out/x.ncu-rep <pixels-per-launch> [kernel-substring]
"""
import collections
import csv
import io
import subprocess
import sys

rep, npx = sys.argv[1], float(sys.argv[2])
kname = sys.argv[3] if len(sys.argv) > 3 else "front"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
segs, cur, hdr, use = [], None, None, False
for r in rows:
    if r and r[0] == "Kernel Name":
        use = kname in r[1] and not segs
        continue
    if r and r[0] == "Address":
        hdr = r
        iS, iN, iI = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        if use:
            stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            iW, iWi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
            cur = {"ops": collections.Counter(), "inst": 0, "samples": 0, "n": 0, "stalls": collections.Counter(), "wf": 0, "wfi": 0}
            segs.append(cur)
        continue
    if not use or hdr is None or len(r) <= iI:
        continue
    sass = r[iS].strip()
    tok = sass.split()
    if not tok:
        continue
    op = tok[1] if tok[0].startswith("@") else tok[0]
    n = int(r[iI] or 0)
    cur["ops"][op.split(".")[0]] += n
    cur["inst"] += n
    cur["samples"] += int(r[iN] or 0)
    cur["n"] += 1
    for i, h in stall_cols:
        cur["stalls"][h[6:]] += int(r[i] or 0)
    cur["wf"] += int(r[iW] or 0)
    cur["wfi"] += int(r[iWi] or 0)
    if op.startswith("BAR"):
        cur = {"ops": collections.Counter(), "inst": 0, "samples": 0, "n": 0, "stalls": collections.Counter(), "wf": 0, "wfi": 0}
        segs.append(cur)
tot = sum(s["inst"] for s in segs)
stot = sum(s["samples"] for s in segs)
print(f"total warp-instr {tot}  = {tot * 32 / npx:.1f} lane-instr/px; samples {stot}")
for i, s in enumerate(segs):
    if not s["inst"]:
        continue
    top = ", ".join(f"{k} {v * 32 / npx:.1f}" for k, v in s["ops"].most_common(12))
    st = ", ".join(f"{k} {100 * v / max(1, sum(s['stalls'].values())):.0f}%" for k, v in s["stalls"].most_common(6))
    print(f"        stalls: {st};  smem wavefronts {s['wf']} (ideal {s['wfi']})")
    print(f"seg {i:2d}: static {s['n']:5d}  exec {s['inst'] * 32 / npx:6.2f}/px ({100 * s['inst'] / tot:4.1f}%)  samples {100 * s['samples'] / stot:4.1f}%  | {top}")
