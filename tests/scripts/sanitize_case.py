"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): fused path on an odd-sized frame (border strips, ragged
right edge), a batch through both hysteresis families, and a 3-band virtual band run; every result checked against the oracle."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np  # noqa: E402

import canny_edge_b200 as cb  # noqa: E402
from canny_edge_b200 import sharded  # noqa: E402
from oracle.bindings import Oracle  # noqa: E402

o = Oracle()
ctx = cb.Context(0)
bad = 0
for h, w, sigma, kind in ((150, 272, 1.4, 0), (97, 250, 1.4, 1), (130, 256, 5.0, 0), (70, 131, 0.8, 1)):
    img = cb.synth_host(1, h, w, kind=kind, seed=h)[0]
    got = cb.cuda_canny(img, sigma, 20, 60, ctx=ctx)
    bad += int((got != o.canny(img, sigma, 20, 60)).sum())
    got5 = cb.cuda_canny(img, sigma, 20, 60, steps=True, ctx=ctx)
    want5 = o.canny(img, sigma, 20, 60, steps=True)
    bad += sum(int((g != wv).sum()) for g, wv in zip(got5, want5))
for kind in (1, 1, 0):
    frames = cb.synth_host(3, 140, 256, kind=kind, seed=5)
    out = cb.canny_batch_host(frames, 1.4, 20, 60, ctx=ctx)
    bad += sum(int((out[f].astype(np.int16) != o.canny(frames[f], 1.4, 20, 60)).sum()) for f in range(3))
img = cb.synth_host(1, 192, 256, kind=1, seed=9)[0]
bad += int((sharded.canny_bands_virtual(img, 3, 1.4, 20, 60) != o.canny(img, 1.4, 20, 60).astype(np.uint8)).sum())
ctx.close()
print("sanitize_case: differing pixels =", bad)
sys.exit(1 if bad else 0)
