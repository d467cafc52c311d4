"""Device-resident interleaved B,G,R frames through b200_canny_batch_device_bgr: Gpix/s with the conversion fused into the front
kernel's staging (default) or as a separate pass (B200_CANNY_BGR_FUSED=0), next to the gray-input rate of the same frames.
    python tools/bgr_probe.py [--frames 64] [--height 2160] [--width 3840] [--sigma 1.4] [--steps 10]
"""
import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import canny_edge_b200 as cb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--sigma", type=float, default=1.4)
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
n, h, w = a.frames, a.height, a.width
ctx = cb.Context(0)
gray = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
cb.load().b200_synth_device(ctx.handle, gray.data_ptr(), n, h, w, 0, 1234, 0)
ctx.synchronize()
# colour frames whose gray value is close to the synthetic frame: B = g - 9, G = g + 3, R = g - 4 (clamped)
g16 = gray.to(torch.int16)
bgr = torch.stack([(g16 - 9).clamp(0, 255), (g16 + 3).clamp(0, 255), (g16 - 4).clamp(0, 255)], dim=-1).to(torch.uint8).contiguous()
del g16
gray2 = torch.empty_like(gray)
cb.load().b200_bgr_to_gray_device(ctx.handle, bgr.data_ptr(), n * h * w, gray2.data_ptr())
out = torch.empty_like(gray)
ref = torch.empty_like(gray)
ctx.synchronize()
torch.cuda.synchronize()
stream = torch.cuda.Stream()          # an explicit stream: handle 0 would mean "the context's own stream" to the library
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)


def timed(fn):
    for _ in range(3):
        fn()
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = stream
    e0.record(st)
    for _ in range(a.steps):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    return ms, n * h * w / ms / 1e6


ms_g, r_g = timed(lambda: cb.canny_batch_device_ptr(ctx, gray2.data_ptr(), n, h, w, a.sigma, 20, 60, ref.data_ptr()))
ms_b, r_b = timed(lambda: cb.canny_batch_device_bgr_ptr(ctx, bgr.data_ptr(), n, h, w, a.sigma, 20, 60, out.data_ptr()))
same = bool(torch.equal(out, ref))
print(json.dumps({"frames": n, "height": h, "width": w, "sigma": a.sigma, "bgr_fused_env": os.environ.get("B200_CANNY_BGR_FUSED", "1"),
                  "gray_input": {"ms": round(ms_g, 4), "gpix_s": round(r_g, 2)},
                  "bgr_input": {"ms": round(ms_b, 4), "gpix_s": round(r_b, 2)},
                  "maps_equal": same, "edge_px": int((out == 255).sum())}))
