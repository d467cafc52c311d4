"""Device-resident batch throughput against the chunk size, for short and long batches (strong scaling leaves 64 frames per GPU
at 8 GPUs).   python tools/chunk_sweep.py [--frames 64,128,512] [--chunks 0,4,5,6,7,8,9,10,12,16]
"""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import canny_edge_b200 as cb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", default="64,128,512")
ap.add_argument("--chunks", default="0,4,5,6,7,8,9,10,12,16")
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--kind", type=int, default=0)
a = ap.parse_args()
h, w = a.height, a.width
frames = [int(x) for x in a.frames.split(",")]
nmax = max(frames)
ctx = cb.Context(0)
d_in = torch.empty((nmax, h, w), dtype=torch.uint8, device="cuda")
d_out = torch.empty_like(d_in)
cb.load().b200_synth_device(ctx.handle, d_in.data_ptr(), nmax, h, w, a.kind, 1234, 0)
ctx.synchronize()
st = torch.cuda.Stream()              # an explicit stream: handle 0 would mean "the context's own stream" to the library
torch.cuda.set_stream(st)
ctx.set_stream(st.cuda_stream)
res = {}
for n in frames:
    for c in [int(x) for x in a.chunks.split(",")]:
        ctx.set_chunk_frames(c)
        fn = lambda: cb.canny_batch_device_ptr(ctx, d_in.data_ptr(), n, h, w, 1.4, 20, 60, d_out.data_ptr())
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = max(2, 1024 // n)
            e0.record(st)
            for _ in range(reps):
                fn()
            e1.record(st)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps)
        res[f"{n}f_chunk{c}"] = {"ms": round(best, 4), "gpix_s": round(n * h * w / best / 1e6, 1)}
        print(f"{n} frames chunk {c}: {best:.4f} ms  {n * h * w / best / 1e6:.1f} Gpix/s", flush=True)
print(json.dumps(res))
