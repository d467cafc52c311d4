"""Pinned-memory PCIe rates on the GPU box: H2D alone, D2H alone, both at once (two streams)."""
import json
import torch

n = 2 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


def chunked_both(chunk=64 << 20):
    for o in range(0, n, chunk):
        with torch.cuda.stream(s1):
            d_a[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
        with torch.cuda.stream(s2):
            h_out[o:o + chunk].copy_(d_b[o:o + chunk], non_blocking=True)


res = {}
for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both), ("both_64MB_chunks", chunked_both)):
    ms = timed(fn)
    res[name] = {"ms": round(ms, 2), "GB_s_per_direction": round(n / ms / 1e6, 2)}
print(json.dumps(res))
