"""Fuzz run of the fused BGR -> gray front kernel (b200_canny_batch_device_bgr) against OpenCV's fixed-point gray formula + the
oracle: random sizes (widths that take the fused kernel and widths that cannot), every half-window, random thresholds, colour noise /
smooth / blocky content, 1-3 frames.    python tests/scripts/fuzz_bgr.py <first_seed> <last_seed+1>
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import canny_edge_b200 as cb  # noqa: E402
from oracle.bindings import Oracle  # noqa: E402


def gray_formula(bgr):
    b, g, r = (bgr[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


n0, n1 = int(sys.argv[1]), int(sys.argv[2])
ctx, oracle = cb.Context(0), Oracle()
cases = 0
for seed in range(n0, n1):
    rng = np.random.default_rng(7000 + seed)
    for case in range(20):
        n = int(rng.integers(1, 4))
        h = int(rng.integers(2, 300))
        w = 16 * int(rng.integers(1, 40)) if rng.random() < 0.7 else int(rng.integers(2, 640))
        sigma = float(rng.choice([0.5, 1.0, 1.4, 1.4, 2.0, 3.0, 5.0, 0.8]))
        lo = int(rng.integers(1, 120))
        hi = int(rng.integers(lo + 1, 256))
        kind = int(rng.integers(0, 3))
        if kind == 0:
            frames = rng.integers(0, 256, (n, h, w, 3)).astype(np.uint8)
        elif kind == 1:
            base = cb.synth_host(n, h, w, kind=0, seed=int(rng.integers(1 << 30))).astype(np.int16)
            frames = np.stack([np.clip(base + rng.integers(-30, 31, (n, h, w)), 0, 255) for _ in range(3)], axis=-1).astype(np.uint8)
        else:
            frames = np.repeat(np.repeat(rng.integers(0, 256, (n, h // 8 + 1, w // 8 + 1, 3)), 8, 1), 8, 2)[:, :h, :w].astype(np.uint8)
        frames = np.ascontiguousarray(frames)
        d_in = torch.from_numpy(frames).cuda()
        d_out = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        cb.canny_batch_device_bgr_ptr(ctx, d_in.data_ptr(), n, h, w, sigma, lo, hi, d_out.data_ptr())
        ctx.synchronize()
        got = d_out.cpu().numpy()
        for f in range(n):
            want = oracle.canny(gray_formula(frames[f]), sigma, lo, hi)
            bad = int((got[f].astype(np.int16) != want).sum())
            assert bad == 0, f"seed {seed} case {case}: {n}x{h}x{w} sigma={sigma} {lo}/{hi} kind={kind} frame {f}: {bad} px differ"
        cases += 1
print(f"bgr fuzz seeds {n0}..{n1 - 1}: {cases} cases, all bit-exact")
