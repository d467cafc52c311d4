"""e2e (pinned host in -> pinned host out) throughput of b200_canny_batch_host vs chunk size and unpack threads."""
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import canny_edge_b200 as cb  # noqa: E402
from canny_edge_b200._lib import check, load  # noqa: E402

n, h, w = int(os.environ.get("FRAMES", "256")), 2160, 3840
lib = load()
ctx = cb.Context(0)
d = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
check(lib.b200_synth_device(ctx.handle, d.data_ptr(), n, h, w, 0, 1234, 0))
torch.cuda.synchronize()
h_in = torch.empty((n, h, w), dtype=torch.uint8, pin_memory=True)
h_out = torch.empty((n, h, w), dtype=torch.uint8, pin_memory=True)
h_in.copy_(d)
del d
res = {}
for chunk in [int(c) for c in os.environ.get("CHUNKS", "2,4,8,16,32").split(",")]:
    ctx.set_chunk_frames(chunk)
    ts = []
    for rep in range(4):
        t0 = time.perf_counter()
        check(lib.b200_canny_batch_host(ctx.handle, h_in.data_ptr(), n, h, w, C.c_float(1.4), 20, 60, h_out.data_ptr()))
        ts.append(time.perf_counter() - t0)
    res[chunk] = round(n * h * w / min(ts[1:]) / 1e9, 2)
print(json.dumps({"frames": n, "Gpix_s_by_chunk_frames": res}))
