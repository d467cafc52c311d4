"""Row-band sharding on ONE GPU: the band kernels (halo-aware front kernel, band-local labelling, boundary
record export, cross-band union, finalisation) run for G virtual bands and must reproduce the oracle's
whole-image result bit for bit — including components that snake across several band boundaries and the
reference's missing (1,0)->(0,1) link (src/utils.cpp:399)."""
import numpy as np
import pytest

import canny_edge_b200 as cb
from canny_edge_b200 import sharded

pytestmark = pytest.mark.gpu


def check(img, bands, sigma, lo, hi, oracle, steps=1, want=None):
    got = sharded.canny_bands_virtual(img, bands, sigma, lo, hi, steps=steps)
    if want is None:
        want = oracle.canny(img, sigma, lo, hi).astype(np.uint8)
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"{img.shape} bands={bands} sigma={sigma} {lo}/{hi}: {len(bad)} px differ, first {bad[0].tolist()}"
    return want


@pytest.mark.parametrize("bands", [1, 2, 3, 8])
def test_bands_shapes_and_noise(oracle, bands):
    for h, w, kind, sigma, lo, hi in [(256, 256, 0, 1.4, 20, 60), (256, 320, 1, 1.4, 20, 60), (257, 333, 1, 1.4, 30, 140),
                                      (400, 128, 1, 1.0, 10, 200), (300, 200, 0, 2.0, 5, 30)]:
        img = cb.synth_host(1, h, w, kind=kind, seed=11 + bands)[0]
        check(img, bands, sigma, lo, hi, oracle)


def test_bands_wide_sigma(oracle):
    img = cb.synth_host(1, 320, 256, kind=1, seed=3)[0]
    check(img, 4, 5.0, 4, 12, oracle)      # 17-row halos, bands of 80 rows
    check(img, 2, 5.0, 4, 12, oracle)


def test_bands_long_weak_chains_across_boundaries(oracle):
    # vertical ramps: long weak edges crossing every band boundary, a single strong seed each
    h, w = 512, 256
    img = np.full((h, w), 100, np.uint8)
    for k, x in enumerate(range(20, w - 20, 24)):
        img[:, x:x + 3] = 108                      # faint vertical stripe (weak edge along the whole height)
        y = (37 * k) % (h - 8)
        img[y:y + 6, x:x + 3] = 200                # one strong blob somewhere on it
    for bands in (2, 4, 8, 16):
        check(img, bands, 1.0, 3, 40, oracle)


def test_bands_quirk_pixel_band_zero(oracle):
    rng = np.random.default_rng(8)
    for _ in range(6):
        img = rng.integers(0, 256, (96, 64)).astype(np.uint8)
        img[:3, :3] = rng.integers(0, 256, (3, 3))
        check(img, 3, 0.8, 10, 250, oracle)


def test_bands_4k_image(oracle):
    img = cb.synth_host(1, 2160, 3840, kind=0, seed=1234)[0]
    want = check(img, 8, 1.4, 20, 60, oracle, steps=3)   # three steps: both parities of the record buffers, every flag reused
    check(img, 5, 1.4, 20, 60, oracle, want=want)        # uneven bands (432 rows each, not a multiple of the slab)


def test_bands_torch_pipeline_agrees(oracle):
    """The round-1 pipeline (exchanges in Python around b200_band_front / _boundary_export / _finalize) and the C handle give the
    same map."""
    img = cb.synth_host(1, 300, 333, kind=1, seed=21)[0]
    want = oracle.canny(img, 1.4, 20, 60).astype(np.uint8)
    assert (sharded.canny_bands_virtual_torch(img, 4, 1.4, 20, 60) == want).all()
    assert (sharded.canny_bands_virtual(img, 4, 1.4, 20, 60) == want).all()


def test_bands_large_image_sigma5_and_gigapixel_class(oracle):
    """Full-size configs inside the driver-run suite: BASELINE configs[3] (8192 x 8192, sigma = 5: 31 taps, 17-row halos) whole and
    in 8 bands, and a 32768-wide band case of configs[4]'s geometry (32768 x 1024 in 8 bands of 128 rows)."""
    img = cb.synth_host(1, 8192, 8192, kind=0, seed=1234)[0]
    want = oracle.canny(img, 5.0, 20, 60).astype(np.uint8)      # ~13 s on one host core
    got = cb.cuda_canny(img, 5.0, 20, 60).astype(np.uint8)
    assert int((got != want).sum()) == 0
    check(img, 8, 5.0, 20, 60, oracle, want=want)
    del img, want, got
    wide = cb.synth_host(1, 1024, 32768, kind=0, seed=99)[0]
    check(wide, 8, 1.4, 20, 60, oracle)


def test_bands_over_nccl_two_gpus():
    """Real 2-rank NCCL run (halo send/recv + record all-gather) when the box has >= 2 GPUs."""
    import subprocess
    import sys
    from pathlib import Path

    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = Path(__file__).resolve().parent / "scripts" / "multigpu_bands_check.py"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script), "--height", "2048", "--width", "2048", "--kind", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"check": "bands_all", "ok": true' in r.stdout
