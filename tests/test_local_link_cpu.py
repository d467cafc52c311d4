"""CPU check of the tile-local hysteresis linking (canny_edge_b200/csrc/local_link.cuh, experimental): the header's
__host__ __device__ functions are driven by a sequential emulation (tests/cpp/local_link_emul.cpp) and the resulting edge maps
are compared with the oracle's hysteresis on random class maps, long chains, tile-border cases and the (0,1)/(1,0) rule."""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
LO, HI = 20, 60


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = tmp_path_factory.mktemp("ll") / "libll_emul.so"
    subprocess.run(["g++", "-std=c++14", "-O2", "-pthread", "-fPIC", "-shared", "-x", "c++", f"-I{ROOT / 'canny_edge_b200' / 'csrc'}",
                    str(ROOT / "tests" / "cpp" / "local_link_emul.cpp"), "-o", str(so)], check=True, capture_output=True, text=True)
    lib = C.CDLL(str(so))
    lib.ll_emulate.restype = C.c_longlong
    lib.ll_emulate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.ll_emulate_mt.restype = C.c_longlong
    lib.ll_emulate_mt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    return lib


def run(emul, oracle, cls, slab_rows=64, threads=1):
    cls = np.ascontiguousarray(cls, np.uint8)
    h, w = cls.shape
    out = np.empty_like(cls)
    visited = emul.ll_emulate_mt(cls.ctypes.data, h, w, slab_rows, 0, out.ctypes.data, threads)
    nms = np.where(cls == 255, HI, np.where(cls == 1, LO, 0)).astype(np.int16)
    want = oracle.hysteresis(nms, LO, HI)
    bad = np.argwhere(out.astype(np.int16) != want)
    assert len(bad) == 0, f"{len(bad)} differing pixels, first {bad[:5].tolist()}"
    return visited, int((cls == 1).sum())


def test_random_maps(emul, oracle):
    rng = np.random.default_rng(5)
    for h, w in ((5, 5), (64, 124), (65, 125), (70, 300), (200, 260), (129, 373), (3, 700)):
        for p_weak, p_strong in ((0.3, 0.01), (0.5, 0.002), (0.1, 0.1), (0.7, 0.0005), (0.45, 0.0)):
            r = rng.random((h, w))
            cls = np.where(r < p_strong, 255, np.where(r < p_strong + p_weak, 1, 0))
            for slab in (64, 17):
                run(emul, oracle, cls, slab)


def test_the_one_way_link(emul, oracle):
    for top in ([0, 1, 255], [0, 1, 0], [255, 1, 0], [1, 1, 255], [0, 255, 0]):
        for left in ([1, 0], [1, 255], [1, 1]):
            cls = np.zeros((4, 4), np.uint8)
            cls[0, :3] = top
            cls[1, 0], cls[2, 0] = left
            run(emul, oracle, cls)
            cls[1, 1] = 1
            run(emul, oracle, cls)


def test_long_chains_and_tile_borders(emul, oracle):
    h, w = 200, 400
    cls = np.zeros((h, w), np.uint8)
    # a serpentine of weak pixels crossing every tile border, one strong pixel at its far end
    for i, y in enumerate(range(2, h - 2, 4)):
        cls[y, 2:w - 2] = 1
        cls[y:y + 5, (w - 3) if i % 2 == 0 else 2] = 1
    run(emul, oracle, cls)            # no seed: nothing survives
    cls[2, 2] = 255
    run(emul, oracle, cls)
    # diagonals through tile corners (rows 63/64, columns 123/124 and 31/32 word borders)
    cls = np.zeros((130, 260), np.uint8)
    for d in range(-60, 60):
        for cx in (124, 32, 96, 248):
            y, x = 64 + d, cx + d
            if 0 <= y < 130 and 0 <= x < 260:
                cls[y, x] = 1
            y, x = 64 + d, cx - 1 - d
            if 0 <= y < 130 and 0 <= x < 260:
                cls[y, x] = 1
    cls[4, 64] = 255
    run(emul, oracle, cls)


def test_only_border_pixels_reach_the_global_kernel(emul, oracle):
    rng = np.random.default_rng(9)
    r = rng.random((640, 1240))
    cls = np.where(r < 0.002, 255, np.where(r < 0.05, 1, 0))
    visited, weak = run(emul, oracle, cls)
    assert visited < 0.08 * weak      # (2*124 + 2*62) / (64*124) = 4.7 % of the tile


def test_concurrent_linking(emul, oracle):
    """The same phases on 8 host threads with real atomics: dense maps with large components, repeated so that different
    interleavings of the lock-free unions occur."""
    rng = np.random.default_rng(21)
    for rep in range(6):
        r = rng.random((128, 248))
        cls = np.where(r < 0.003, 255, np.where(r < 0.55, 1, 0))
        run(emul, oracle, cls, threads=8)
