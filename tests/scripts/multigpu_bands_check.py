"""Run under torchrun on N >= 2 GPUs: row-band sharded Canny of one synthetic image over NCCL, gathered on
rank 0 and compared bit for bit with the CPU oracle on the whole image.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/scripts/multigpu_bands_check.py --height 4096 --width 4096 --kind 1
"""
import argparse
import ctypes as C
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import canny_edge_b200 as cb  # noqa: E402
from canny_edge_b200 import sharded  # noqa: E402
from canny_edge_b200._lib import check, load  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=4096)
    ap.add_argument("--width", type=int, default=4096)
    ap.add_argument("--kind", type=int, default=1)
    ap.add_argument("--sigma", type=float, default=1.4)
    ap.add_argument("--lo", type=int, default=20)
    ap.add_argument("--hi", type=int, default=60)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = load()
    ctx = cb.Context(local)
    # an explicit stream: torch's DEFAULT stream is handle 0, which b200_ctx_set_stream takes as "use the context's private stream" —
    # the context's kernels would then not be ordered with torch's own work (zero_, NCCL) on the default stream
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sizes = [sharded.band_geometry(a.height, a.width, r, world, a.sigma).rows for r in range(world)]
    want = None
    if rank == 0:
        from oracle.bindings import Oracle
        img = cb.synth_host(1, a.height, a.width, kind=a.kind, seed=1234)[0]
        want = Oracle().canny(img, a.sigma, a.lo, a.hi).astype(np.uint8)
    ok = True
    # the C pipeline over both transports (P2P peer mappings, NCCL), two steps each (the second reuses every buffer and flag), and
    # the round-1 torch.distributed pipeline as a third opinion
    for name in ("p2p", "nccl", "torch"):
        if name == "torch":
            pipe = sharded.TorchBandPipeline(ctx, a.height, a.width, rank, world, a.sigma, a.lo, a.hi)
        else:
            os.environ["B200_BANDS_TRANSPORT"] = name
            pipe = sharded.BandPipeline(ctx, a.height, a.width, rank, world, a.sigma, a.lo, a.hi)
        g = pipe.geo
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        band = pipe.band_view()
        edges = torch.empty((g.rows, g.width), dtype=torch.uint8, device="cuda")
        for step in range(2):
            check(lib.b200_synth_rows_device(ctx.handle, band.data_ptr(), g.row0, g.rows, g.width, a.kind, 1234, 0))
            edges.zero_()
            pipe.run(None, edges)
            if name != "torch":
                pipe.check()
            torch.cuda.synchronize()
            parts = [torch.empty((s, a.width), dtype=torch.uint8, device="cuda") for s in sizes] if rank == 0 else None
            dist.gather(edges, parts, dst=0)
            if rank == 0:
                got = torch.cat(parts).cpu().numpy()
                bad = int((got != want).sum())
                ok = ok and bad == 0
                print(json.dumps({"check": "bands_vs_oracle", "pipeline": name, "transport": getattr(pipe, "transport", None), "step": step,
                                  "world": world, "height": a.height, "width": a.width, "kind": a.kind,
                                  "edge_pixels": int((got == 255).sum()), "differing_pixels": bad, "ok": bad == 0}))
        if name != "torch":
            pipe.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    ok = bool(flag.item())
    if rank == 0:
        print(json.dumps({"check": "bands_all", "ok": ok}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
