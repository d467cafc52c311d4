"""Single-frame latency configuration (BASELINE configs[1]) with and without programmatic dependent launch between the front kernel
and the two list-driven hysteresis kernels: run once per B200_CANNY_PDL value.   python tools/pdl_probe.py
"""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import canny_edge_b200 as cb  # noqa: E402

ctx = cb.Context(0)
stream = torch.cuda.Stream()          # an explicit stream: handle 0 would mean "the context's own stream" to the library
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
res = {"pdl_env": os.environ.get("B200_CANNY_PDL", "default")}
for name, h, w, kind in (("256x256", 256, 256, 0), ("1080p_shapes", 1080, 1920, 0), ("1080p_noise", 1080, 1920, 1), ("4k_shapes", 2160, 3840, 0)):
    d_in = torch.empty((1, h, w), dtype=torch.uint8, device="cuda")
    d_out = torch.empty_like(d_in)
    cb.load().b200_synth_device(ctx.handle, d_in.data_ptr(), 1, h, w, kind, 1234, 0)
    ctx.synchronize()
    st = stream
    fn = lambda: cb.canny_batch_device_ptr(ctx, d_in.data_ptr(), 1, h, w, 1.4, 20, 60, d_out.data_ptr())
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(200):
            fn()
        e1.record(st)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 200 * 1e3)
    res[name] = {"us_per_frame": round(best, 2), "edge_px": int((d_out == 255).sum()),
                 "checksum": int(d_out.view(-1).to(torch.int64).mul(torch.arange(h * w, device="cuda") % 65521).sum().item())}
print(json.dumps(res))
