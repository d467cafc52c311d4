"""BASELINE configs[1] and [3] through the public API: one 1920x1080 frame (latency path, host buffers, what cuda_canny does per
frame) and one 8192x8192 image at sigma=5 (31-tap window), plus the device-resident times of the same calls.

    python tools/latency_probe.py
"""
import ctypes as C
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import canny_edge_b200 as cb  # noqa: E402
from canny_edge_b200._lib import check, load  # noqa: E402

lib = load()
ctx = cb.Context(0)
out = {}
for name, h, w, sigma in (("1080p_sigma1.4", 1080, 1920, 1.4), ("4k_sigma1.4", 2160, 3840, 1.4), ("8192sq_sigma5", 8192, 8192, 5.0)):
    img = cb.synth_host(1, h, w, kind=0, seed=1234)[0]
    edges = np.empty((h, w), np.int16)
    e8 = np.empty((1, h, w), np.uint8)

    def host_call():
        check(lib.b200_canny(ctx.handle, img.ctypes.data, C.c_float(sigma), 20, 60, h, w, edges.ctypes.data))

    def host_u8_call():
        check(lib.b200_canny_batch_host(ctx.handle, img.ctypes.data, 1, h, w, C.c_float(sigma), 20, 60, e8.ctypes.data))

    # the same two calls on pinned host buffers (what a capture pipeline that allocates its frames through the library gets)
    pin = [C.c_void_p() for _ in range(3)]
    for p_, nbytes in zip(pin, (h * w, h * w, 2 * h * w)):
        check(lib.b200_host_alloc_pinned(nbytes, C.byref(p_)))
    C.memmove(pin[0], img.ctypes.data, h * w)

    def host_u8_pinned():
        check(lib.b200_canny_batch_host(ctx.handle, pin[0], 1, h, w, C.c_float(sigma), 20, 60, pin[1]))

    def host_i16_pinned():
        check(lib.b200_canny(ctx.handle, pin[0], C.c_float(sigma), 20, 60, h, w, pin[2]))

    res = {}
    for label, fn in (("b200_canny_host_i16_ms", host_call), ("b200_canny_batch_host_u8_ms", host_u8_call),
                      ("b200_canny_host_i16_pinned_ms", host_i16_pinned), ("b200_canny_batch_host_u8_pinned_ms", host_u8_pinned)):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            t0 = time.perf_counter()
            fn()
            ts.append((time.perf_counter() - t0) * 1e3)
        res[label] = round(sorted(ts)[len(ts) // 2], 3)
    d_in = torch.from_numpy(img).cuda()
    d_out = torch.empty_like(d_in)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    for _ in range(3):
        cb.canny_batch_device_ptr(ctx, d_in.data_ptr(), 1, h, w, sigma, 20, 60, d_out.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        cb.canny_batch_device_ptr(ctx, d_in.data_ptr(), 1, h, w, sigma, 20, 60, d_out.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    res["device_resident_ms"] = round(e0.elapsed_time(e1) / 20, 4)
    res["device_resident_Mpix_s"] = round(h * w / res["device_resident_ms"] / 1e3, 1)
    ctx.set_stream(0)
    torch.cuda.set_stream(torch.cuda.default_stream())
    for p_ in pin:
        check(lib.b200_host_free_pinned(p_))
    out[name] = res
print(json.dumps(out))
