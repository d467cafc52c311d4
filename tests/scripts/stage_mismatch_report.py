"""Debug aid: runs one synthetic frame through b200_canny_steps and the oracle and reports, per stage plane, where they differ.

    python tests/scripts/stage_mismatch_report.py HEIGHT WIDTH SIGMA
"""
import sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np
import canny_edge_b200 as cb
from oracle.bindings import Oracle
o = Oracle()
ctx = cb.Context(0)
h, w = int(sys.argv[1]), int(sys.argv[2]); sigma = float(sys.argv[3])
img = cb.synth_host(1, h, w, kind=0, seed=3)[0]
try:
    got = cb.cuda_canny(img, sigma, 20, 60, steps=True, ctx=ctx)
except Exception as e:
    print("ERROR", e); sys.exit(0)
want = o.canny(img, sigma, 20, 60, steps=True)
for n, g, wv in zip(("blur", "mag", "ang", "nms", "edges"), got, want):
    bad = np.argwhere(g != wv)
    print(n, "mismatches", len(bad), bad[:5].tolist(), [(int(g[tuple(b)]), int(wv[tuple(b)])) for b in bad[:5]])
