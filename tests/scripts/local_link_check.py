"""EXPERIMENTAL path check (run by hand on a B200, under a timeout):

    B200_CANNY_LOCAL_LINK=1 timeout 120 python tests/scripts/local_link_check.py

Runs frames through the device batch path with tile-local hysteresis linking enabled in the front kernel (front2.cu LL,
local_link.cuh) and compares every edge map with the oracle; then times it against the regular path.  Not part of the pytest
suite: the path has not been validated on hardware yet and a bug in its shared-memory union-find could hang the GPU."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import canny_edge_b200 as cb  # noqa: E402
from oracle.bindings import Oracle  # noqa: E402

assert os.environ.get("B200_CANNY_LOCAL_LINK") == "1", "set B200_CANNY_LOCAL_LINK=1 (read once by the library)"
oracle, ctx = Oracle(), cb.Context(0)
bad = 0
for n, h, w, kind in ((2, 270, 480, 0), (3, 301, 333, 1), (2, 1080, 1920, 0), (2, 1080, 1920, 1), (1, 2160, 3840, 0), (4, 64, 124, 1),
                      (2, 130, 250, 1), (1, 65, 125, 1)):
    frames = cb.synth_host(n, h, w, kind=kind, seed=100 + h)
    d_in = torch.from_numpy(frames).cuda()
    d_out = torch.empty_like(d_in)
    torch.cuda.synchronize()
    cb.canny_batch_device_ptr(ctx, d_in.data_ptr(), n, h, w, 1.4, 20, 60, d_out.data_ptr())
    ctx.synchronize()
    got = d_out.cpu().numpy()
    for f in range(n):
        diff = int((got[f].astype(np.int16) != oracle.canny(frames[f], 1.4, 20, 60)).sum())
        bad += diff
        print(f"{n}x{h}x{w} kind {kind} frame {f}: {diff} differing pixels")
print("local-link path:", "bit-exact" if bad == 0 else f"{bad} differing pixels")
sys.exit(0 if bad == 0 else 1)
