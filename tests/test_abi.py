"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, its host-side helpers match the oracle, and compute entry points fail LOUDLY (no CPU
fallback) when no B200 is present."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import canny_edge_b200 as cb
from canny_edge_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "canny_b200.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/canny_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in canny_edge_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table lists symbols the header does not declare"


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.b200_version() == 1
    assert isinstance(lib.b200_last_error(), bytes)


def test_host_gaussian_kernel_matches_oracle(oracle):
    for sigma in (0.3, 0.5, 0.8, 1.0, 1.4, 2.0, 2.5, 3.0, 5.0, 7.0, 16.0):
        w, n = cb.createGaussianKernel(sigma)
        wo, no = oracle.gaussian_kernel(sigma)
        assert n == no and (w == wo).all(), sigma
        assert (w == w[::-1]).all()  # the symmetry the fused blur's product sharing relies on


def test_host_direction_table_matches_oracle(oracle):
    want = oracle.angle_table(1020)
    got = np.empty_like(want)
    assert _lib.load().b200_direction_table_host(1020, got.ctypes.data) == 0
    assert int((got != want).sum()) == 0  # all 4,165,681 (gx,gy) pairs a blurred 8-bit image can produce


def test_host_isqrt_matches_oracle(oracle):
    lib = _lib.load()
    tab = oracle.isqrt_table(2 * 1020 * 1020)
    idx = np.unique(np.concatenate([np.arange(0, 5000), np.arange(1, 1443) ** 2, np.arange(1, 1443) ** 2 - 1,
                                    np.random.default_rng(0).integers(0, 2 * 1020 * 1020, 20000)]))
    idx = idx[idx <= 2 * 1020 * 1020]
    for n in idx:
        assert lib.b200_isqrt_host(int(n)) == tab[n]


def test_synth_host_is_deterministic_and_row_addressable():
    a = cb.synth_host(2, 40, 70, kind=0, seed=1234, first_frame=3)
    b = cb.synth_rows_host(10, 20, 70, kind=0, seed=1234, frame=4)
    assert (a[1, 10:30] == b).all()
    assert (cb.synth_host(1, 8, 8, kind=2)[0] == 128).all()
    assert len(np.unique(cb.synth_host(1, 64, 64, kind=1)[0])) > 100


def test_band_geometry_helpers():
    lib = _lib.load()
    assert lib.b200_band_halo_rows(C.c_float(1.4)) == 7   # window/2 + Sobel + NMS
    assert lib.b200_band_halo_rows(C.c_float(5.0)) == 17
    assert lib.b200_band_record_count(100) == 202


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(cb.CannyB200Error) as ei:
        cb.cuda_canny(np.zeros((8, 8), np.uint8), 1.4, 20, 60)
    assert ei.value.status == _lib.ERR_NO_DEVICE
    with pytest.raises(cb.CannyB200Error):
        cb.Context(0)


def test_argument_validation_without_device():
    lib = _lib.load()
    assert lib.b200_gaussian_kernel(C.c_float(-1.0), None, None) == _lib.ERR_INVALID_ARG
    assert lib.b200_gaussian(None, None, C.c_float(1.4), 4, 4, None) == _lib.ERR_INVALID_ARG
    z = np.zeros((1, 8), np.uint8)
    o = np.zeros((1, 8), np.int16)
    # height < 2 is rejected before any device work (the reference reads out of bounds there)
    assert lib.b200_gaussian(None, z.ctypes.data, C.c_float(1.4), 1, 8, o.ctypes.data) == _lib.ERR_INVALID_ARG


def test_unpack_edges_host_matches_numpy():
    """The host half of the bit-packed edge transfer (pure host code, thread pool + streaming stores): against numpy's
    unpackbits for byte and int16 output, one and several threads, aligned and misaligned outputs, pixel counts that are not
    multiples of 8 / 64, sizes on both sides of the single-thread shortcut."""
    lib = _lib.load()
    rng = np.random.default_rng(11)
    for n_px in (1, 7, 8, 63, 64, 1000, 4097, 300_001, 2_000_003):
        bits = rng.integers(0, 256, (n_px + 7) // 8, dtype=np.uint8)
        want = np.unpackbits(bits, bitorder="little")[:n_px].astype(np.int16) * 255
        for threads in (1, 4, 0):
            for elem, dt in ((1, np.uint8), (2, np.int16)):
                for off in (0, 1):
                    raw = np.full(n_px + 1 + 8, 77, dt)
                    out = raw[off:off + n_px]
                    assert lib.b200_unpack_edges_host(bits.ctypes.data, n_px, out.ctypes.data, elem, threads) == 0
                    assert (out.astype(np.int16) == want).all(), (n_px, threads, elem, off)
                    assert (raw[off + n_px:] == 77).all() and (raw[:off] == 77).all(), "wrote outside the output"
    assert lib.b200_unpack_edges_host(None, 8, None, 1, 1) != 0
    assert lib.b200_unpack_edges_host(bits.ctypes.data, 8, out.ctypes.data, 4, 1) != 0


def test_host_bind_numa_is_harmless_without_gpu():
    """b200_host_bind_numa never fails the caller: without a device it reports an error status and changes nothing."""
    import ctypes as C
    import os
    lib = cb.load()
    before = os.sched_getaffinity(0)
    node = C.c_int(7)
    st = lib.b200_host_bind_numa(0, C.byref(node))
    assert st in (0, 2)                     # B200_OK on a GPU box, B200_ERR_NO_DEVICE here
    assert os.sched_getaffinity(0) <= before   # only ever narrowed
    os.sched_setaffinity(0, before)
