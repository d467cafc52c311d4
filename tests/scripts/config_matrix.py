"""Config x input-family matrix of SURVEY 8(d): every BASELINE.json config (C1..C5) on the four synthetic input families, the
CUDA path timed device-resident (CUDA events on the launching stream) and checked bit for bit against the compiled reference
(oracle/_ref, single thread, timed with the reference's own std::chrono bracket, src/utils.cpp:435,479) wherever the reference
finishes in reasonable time:

    C1 256x256 (full), C2 1920x1080 (full), C3 512x3840x2160 (GPU: all 512 frames; reference: first 4), C4 8192x8192 sigma 5 (full),
    C5 32768x32768 (full image through the reference once, `shapes` only: ~10 GB of host memory and a minute or two)

Not collected by pytest (no test_ prefix): it is a measurement script that lives under tests/ because it drives the oracle.

    python tests/scripts/config_matrix.py [--skip-c5-ref] [--out gpurun_out/config_matrix.json]
"""
import argparse
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import canny_edge_b200 as cb  # noqa: E402
from canny_edge_b200._lib import check, load  # noqa: E402
from oracle.bindings import REF_LIB_O0, Ref  # noqa: E402

LO, HI = 20, 60
KINDS = {"shapes": 0, "noise": 1, "const128": 2, "testjpg_tiled": -1}
CONFIGS = [
    # name, frames, H, W, sigma, frames through the reference
    ("C1_256x256", 1, 256, 256, 1.4, 1),
    ("C2_1080p", 1, 1080, 1920, 1.4, 1),
    ("C3_512x4K", 512, 2160, 3840, 1.4, 4),
    ("C4_8192sq_s5", 1, 8192, 8192, 5.0, 1),
    ("C5_32768sq", 1, 32768, 32768, 1.4, 1),
]


def make_input(lib, ctx, kind, n, h, w, tile):
    d = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    if kind >= 0:
        check(lib.b200_synth_device(ctx.handle, d.data_ptr(), n, h, w, kind, 1234, 0))
    else:
        ry, rx = -(-h // tile.shape[0]), -(-w // tile.shape[1])
        d[:] = tile.repeat(ry, rx)[:h, :w]
    torch.cuda.synchronize()
    return d


def gpu_time(ctx, d_in, d_out, sigma, reps):
    n, h, w = d_in.shape
    for _ in range(3):
        cb.canny_batch_device_ptr(ctx, d_in.data_ptr(), n, h, w, sigma, LO, HI, d_out.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        cb.canny_batch_device_ptr(ctx, d_in.data_ptr(), n, h, w, sigma, LO, HI, d_out.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def mix64_np(x):
    """splitmix64 finaliser on a uint64 array (canny_math.h::mix64), wrapping arithmetic."""
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-c5-ref", action="store_true")
    ap.add_argument("--only", default="", help="comma-separated config name prefixes (e.g. C4,C5)")
    ap.add_argument("--out", default="gpurun_out/config_matrix.json")
    a = ap.parse_args()

    lib = load()
    ctx = cb.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ref = Ref()
    ref_o0 = Ref(REF_LIB_O0) if REF_LIB_O0.exists() else None
    tile = torch.from_numpy(np.fromfile(ROOT / "tests" / "golden" / "test_gray_256x256.u8", dtype=np.uint8).reshape(256, 256)).cuda()

    rows = []
    only = tuple(s for s in a.only.split(",") if s)
    for name, n, h, w, sigma, n_ref in CONFIGS:
        if only and not name.startswith(only):
            continue
        for kname, kind in KINDS.items():
            d_in = make_input(lib, ctx, kind, n, h, w, tile)
            d_out = torch.empty_like(d_in)
            px = n * h * w
            ms = gpu_time(ctx, d_in, d_out, sigma, 10 if px < (1 << 28) else 5)
            edges = C.c_ulonglong()
            check(lib.b200_count_edges_device(ctx.handle, d_out.data_ptr(), d_out.numel(), C.byref(edges)))
            row = {"config": name, "input": kname, "frames": n, "height": h, "width": w, "sigma": sigma, "gpu_ms": round(ms, 4),
                   "gpu_Mpix_s": round(px / ms / 1e3, 1), "hbm_frac_2Bpx": round(px * 2 / (ms * 1e-3) / 6449.1e9, 4),
                   "edge_frac": round(edges.value / px, 5)}
            run_ref = not (name.startswith("C5") and (a.skip_c5_ref or kname != "shapes"))
            if run_ref:
                h_in = d_in[:n_ref].cpu().numpy()
                h_out = d_out[:n_ref].cpu().numpy()
                secs, diff = 0.0, 0
                for f in range(n_ref):
                    e = ref.canny(h_in[f], sigma, LO, HI)
                    secs += ref.last_seconds
                    diff += int(np.count_nonzero((e != 0) != (h_out[f] != 0)))
                    if name.startswith("C5"):
                        # edge count + position-dependent checksum of the REFERENCE's map, computed on the host: the pair bench.py
                        # expects from the row-band run at every GPU count (BANDS_EXPECT) is anchored here
                        idx = np.flatnonzero(e.ravel() == 255).astype(np.uint64)
                        row["ref_edge_pixels"] = int(idx.size)
                        row["ref_checksum"] = f"{int(mix64_np(idx).sum(dtype=np.uint64)):016x}"
                        hs = C.c_ulonglong()
                        check(lib.b200_hash_edges_device(ctx.handle, d_out.data_ptr(), d_out.numel(), 0, C.byref(hs)))
                        row["gpu_edge_pixels"], row["gpu_checksum"] = int(edges.value), f"{hs.value:016x}"
                        del idx
                    del e
                row.update({"ref_frames": n_ref, "ref_1thread_s": round(secs, 4), "ref_1thread_Mpix_s": round(n_ref * h * w / secs / 1e6, 2),
                            "differing_px": diff, "speedup_vs_1thread": round((px / ms / 1e3) / (n_ref * h * w / secs / 1e6), 1)})
                if ref_o0 is not None and kname == "shapes" and name.startswith(("C1", "C2")):
                    ref_o0.canny(h_in[0], sigma, LO, HI, want_edges=False)
                    row["ref_O0_1thread_Mpix_s"] = round(h * w / ref_o0.last_seconds / 1e6, 2)
                del h_in, h_out
            rows.append(row)
            print(json.dumps(row), flush=True)
            del d_in, d_out
            torch.cuda.empty_cache()
    out = Path(a.out)
    out.parent.mkdir(parents=True, exist_ok=True)
    out.write_text(json.dumps({"thresholds": [LO, HI], "gpu": torch.cuda.get_device_name(0), "rows": rows,
                               "when": time.strftime("%Y-%m-%d %H:%M:%S")}, indent=1))


if __name__ == "__main__":
    main()
