"""Per-kernel timing of the pipeline on synthetic frames resident in HBM (CUDA events around every
kernel, serial on one stream) + whole-pipeline throughput with the three-stream chunk overlap.

    python tools/stage_times.py --frames 64 --height 2160 --width 3840 --sigma 1.4 --kind 0
"""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import canny_edge_b200 as cb  # noqa: E402
from canny_edge_b200._lib import check, load  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--sigma", type=float, default=1.4)
ap.add_argument("--kind", type=int, default=0)
ap.add_argument("--lo", type=int, default=20)
ap.add_argument("--hi", type=int, default=60)
ap.add_argument("--chunk", type=int, default=0)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()

lib = load()
ctx = cb.Context(0)
stream = torch.cuda.Stream()  # a real (non-default) stream: events below are recorded on the stream the kernels run on
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
ctx.set_chunk_frames(a.chunk)
n, h, w = a.frames, a.height, a.width
d_in = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
d_out = torch.empty_like(d_in)
if a.kind >= 0:
    check(lib.b200_synth_device(ctx.handle, d_in.data_ptr(), n, h, w, a.kind, 1234, 0))
else:
    # --kind -1: the reference's tests/test.jpg (decoded gray, committed as raw bytes) tiled over every frame
    import numpy as np
    tile = torch.from_numpy(np.fromfile(Path(__file__).resolve().parents[1] / "tests" / "golden" / "test_gray_256x256.u8", dtype=np.uint8).reshape(256, 256)).cuda()
    d_in[:] = tile.repeat(-(-h // 256), -(-w // 256))[:h, :w]
torch.cuda.synchronize()
px = n * h * w

ms = (C.c_float * 5)()
cnt = (C.c_int * 5)()
for _ in range(2):
    check(lib.b200_profile_stages_device(ctx.handle, d_in.data_ptr(), n, h, w, C.c_float(a.sigma), a.lo, a.hi, d_out.data_ptr(), ms, cnt))
names = ["front", "ccl_local", "ccl_merge", "ccl_final", "other"]
stage = {names[i]: {"ms": round(ms[i], 4), "launches": cnt[i], "Mpix_s": round(px / ms[i] / 1e3, 1) if ms[i] > 0 else None} for i in range(5)}

for _ in range(3):
    cb.canny_batch_device_ptr(ctx, d_in.data_ptr(), n, h, w, a.sigma, a.lo, a.hi, d_out.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    cb.canny_batch_device_ptr(ctx, d_in.data_ptr(), n, h, w, a.sigma, a.lo, a.hi, d_out.data_ptr())
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / a.reps
edges = C.c_ulonglong()
check(lib.b200_count_edges_device(ctx.handle, d_out.data_ptr(), d_out.numel(), C.byref(edges)))
print(json.dumps({"shape": [n, h, w], "sigma": a.sigma, "kind": a.kind, "chunk": a.chunk, "pipeline_ms": round(t, 4),
                  "pipeline_Mpix_s": round(px / t / 1e3, 1), "roofline_frac_2Bpx_6449GBs": round(px * 2 / (t * 1e-3) / 6449.1e9, 4),
                  "edge_frac": edges.value / px, "stages": stage}))
