"""Generates the golden fixtures from the UNMODIFIED reference (oracle/_ref/libcanny_ref.so, i.e.
/root/reference/src/utils.cpp compiled by oracle/Makefile).  Run in the authoring container only:

    python tests/golden/make_golden.py

Outputs (committed): test_gray_256x256.u8 (decoded reference test image), golden_*.npz, manifest.json.
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.bindings import Ref  # noqa: E402

HERE = Path(__file__).resolve().parent
ref = Ref()
manifest = {}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def stages(img, sigma, lo, hi):
    blur = ref.gaussian(img, sigma)
    mag, ang = ref.sobel(blur)
    nms = ref.nonmaximal(mag, ang)
    edges = ref.hysteresis(nms, lo, hi)
    assert (edges == ref.canny(img, sigma, lo, hi)).all()
    return blur, mag, ang, nms, edges


# 1. the reference's own test image (tests/test.jpg), BASELINE config 1
import cv2  # noqa: E402

gray = cv2.imread("/root/reference/tests/test.jpg", cv2.IMREAD_GRAYSCALE)
assert gray.shape == (256, 256)
gray.tofile(HERE / "test_gray_256x256.u8")
manifest["test_gray_sha256"] = sha(gray)
for sigma, lo, hi in [(1.4, 20, 60), (0.5, 10, 50), (2.0, 20, 60), (5.0, 5, 15)]:
    blur, mag, ang, nms, edges = stages(gray, sigma, lo, hi)
    key = f"testjpg_s{sigma}_{lo}_{hi}"
    manifest[key] = {
        "blur_sha256": sha(blur), "mag_sha256": sha(mag), "ang_sha256": sha(ang), "nms_sha256": sha(nms),
        "edges_sha256": sha(edges), "edge_pixels": int((edges == 255).sum()), "blur_sum": int(blur.sum(dtype=np.int64)),
        "mag_sum": int(mag.sum(dtype=np.int64)), "nms_sum": int(nms.sum(dtype=np.int64)),
        "angle_counts": {str(a): int((ang == a).sum()) for a in (0, 45, 90, 135)},
    }
    if (sigma, lo, hi) == (1.4, 20, 60):
        np.savez_compressed(HERE / "golden_testjpg_s1.4_20_60.npz", blur=blur, mag=mag, ang=ang, nms=nms,
                            edges=(edges == 255).astype(np.uint8))

# 2. small random / structured cases with every plane stored (odd sizes, W < window, all sigmas)
rng = np.random.default_rng(20261018)
cases = {}
shapes = [(2, 2), (2, 9), (9, 2), (3, 3), (5, 7), (17, 4), (31, 33), (64, 64), (65, 127), (40, 130), (70, 251)]
for i, (h, w) in enumerate(shapes):
    for sigma in (0.5, 1.0, 1.4, 2.0, 3.0, 5.0):
        kind = (i + int(sigma * 10)) % 3
        if kind == 0:
            img = rng.integers(0, 256, (h, w)).astype(np.uint8)
        elif kind == 1:
            yy, xx = np.mgrid[0:h, 0:w]
            img = (128 + 100 * np.sin(xx / 3.0) * np.cos(yy / 5.0) + rng.integers(-6, 7, (h, w))).clip(0, 255).astype(np.uint8)
        else:
            img = ((rng.random((h, w)) < 0.25) * rng.integers(60, 256, (h, w))).astype(np.uint8)
        lo = int(rng.integers(0, 80))
        hi = int(rng.integers(lo + 1, 200))
        blur, mag, ang, nms, edges = stages(img, sigma, lo, hi)
        k = f"c{i}_s{sigma}"
        cases[k + "_img"] = img
        cases[k + "_par"] = np.array([lo, hi], np.int32)
        cases[k + "_blur"] = blur
        cases[k + "_mag"] = mag
        cases[k + "_ang"] = ang
        cases[k + "_nms"] = nms
        cases[k + "_edges"] = (edges == 255).astype(np.uint8)
np.savez_compressed(HERE / "golden_small_cases.npz", **cases)
manifest["small_cases"] = len(shapes) * 6

# 3. hysteresis-only cases incl. the (1,0)->(0,1) missing link (src/utils.cpp:399)
hc = {}
for i in range(200):
    h, w = rng.integers(2, 9, 2)
    nms = (rng.integers(0, 4, (h, w)) * 40).astype(np.int16)  # 0 / 40 (weak) / 80, 120 (strong) with lo=20, hi=60
    hc[f"h{i}_in"] = nms
    hc[f"h{i}_out"] = (ref.hysteresis(nms, 20, 60) == 255).astype(np.uint8)
np.savez_compressed(HERE / "golden_hysteresis_cases.npz", **hc)
manifest["hysteresis_cases"] = 200

(HERE / "manifest.json").write_text(json.dumps(manifest, indent=1, sort_keys=True))
print(json.dumps(manifest["testjpg_s1.4_20_60"], indent=1))
print("wrote", [p.name for p in HERE.iterdir()])
