"""TEST INFRASTRUCTURE ONLY: ctypes bindings of the two CPU checkers.

  Oracle  -> oracle/liboracle.so          plain-C restatement (canny_oracle.c)
  Ref     -> oracle/_ref/libcanny_ref.so  the unmodified reference src/utils.cpp behind ref_shim.cpp
             (prebuilt in the authoring container; absent only if it was never built)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_LIB = HERE / "liboracle.so"
REF_LIB = HERE / "_ref" / "libcanny_ref.so"
REF_LIB_O0 = HERE / "_ref" / "libcanny_ref_O0.so"   # same sources at -O0 (what the reference's CMake builds); timing only


def build(verbose: bool = False) -> None:
    """make -C oracle: the C restatement always, _ref only where /root/reference exists."""
    res = subprocess.run(["make", "-C", str(HERE)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stdout)


def _p(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class _Base:
    prefix = ""

    def __init__(self, path: Path):
        if not path.exists():
            raise FileNotFoundError(f"{path} not built (run `make -C oracle`)")
        self.lib = C.CDLL(str(path))
        self.path = path

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    # createGaussianKernel, src/utils.cpp:77-95
    def gaussian_kernel(self, sigma: float):
        w = np.zeros(1024, np.float32)
        n = C.c_int()
        self._f("gaussian_kernel")(C.c_float(sigma), _p(w), C.byref(n))
        return w[: n.value].copy(), n.value

    def gaussian(self, img, sigma: float):
        a = _c(img, np.uint8)
        out = np.empty(a.shape, np.int16)
        self._f("gaussian")(_p(a), C.c_float(sigma), a.shape[0], a.shape[1], _p(out))
        return out

    def xy_gradient(self, blur):
        a = _c(blur, np.int16)
        gx, gy = np.empty(a.shape, np.int16), np.empty(a.shape, np.int16)
        self._f("xy_gradient")(_p(a), a.shape[0], a.shape[1], _p(gx), _p(gy))
        return gx, gy

    def sobel(self, blur):
        a = _c(blur, np.int16)
        m, g = np.empty(a.shape, np.int16), np.empty(a.shape, np.int16)
        self._f("sobel")(_p(a), a.shape[0], a.shape[1], _p(m), _p(g))
        return m, g

    def nonmaximal(self, mag, ang):
        m, a = _c(mag, np.int16), _c(ang, np.int16)
        out = np.empty(m.shape, np.int16)
        self._f("nonmaximal")(_p(m), _p(a), m.shape[0], m.shape[1], _p(out))
        return out

    def hysteresis(self, nms, lo: int, hi: int):
        a = _c(nms, np.int16).copy()
        self._f("hysteresis")(_p(a), a.shape[0], a.shape[1], int(lo), int(hi))
        return a

    def find_edge_pixels(self, nms, visited, start: int, lo: int, hi: int):
        a = _c(nms, np.int16).copy()
        v = _c(visited, np.uint8).copy()
        self._f("find_edge_pixels")(_p(a), _p(v), int(start), int(lo), int(hi), a.shape[0], a.shape[1])
        return a, v


class Oracle(_Base):
    prefix = "oracle_"

    def __init__(self):
        if not ORACLE_LIB.exists():
            build()
        super().__init__(ORACLE_LIB)
        self.lib.oracle_canny.restype = C.c_double

    def angle_table(self, gmax: int):
        n = 2 * gmax + 1
        out = np.empty((n, n), np.int16)
        self.lib.oracle_angle_table(int(gmax), _p(out))
        return out

    def isqrt_table(self, n_max: int):
        out = np.empty(n_max + 1, np.int32)
        self.lib.oracle_isqrt_table(int(n_max), _p(out))
        return out

    def canny(self, img, sigma: float, lo: int, hi: int, steps: bool = False):
        """Returns edges (int16 0/255) or, with steps, (blur, mag, ang, nms, edges); .last_seconds is set."""
        a = _c(img, np.uint8)
        h, w = a.shape
        edges = np.empty((h, w), np.int16)
        if steps:
            planes = [np.empty((h, w), np.int16) for _ in range(4)]
            ptrs = [_p(x) for x in planes]
        else:
            planes, ptrs = [], [None] * 4
        self.last_seconds = self.lib.oracle_canny(_p(a), C.c_float(sigma), int(lo), int(hi), h, w, _p(edges), *ptrs)
        if self.last_seconds < 0:
            raise RuntimeError(f"oracle_canny failed ({self.last_seconds})")
        return (*planes, edges) if steps else edges


def synth_frames(n_frames: int, height: int, width: int, kind: int = 0, seed: int = 1234, first_frame: int = 0, threads: int = 1) -> np.ndarray:
    """The bench's procedural frames from the ORACLE's own generator (oracle_synth_rows): no product code involved."""
    import threading

    if not ORACLE_LIB.exists():
        build()
    lib = C.CDLL(str(ORACLE_LIB))
    lib.oracle_synth_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int]
    out = np.empty((n_frames, height, width), np.uint8)
    jobs = list(range(n_frames))
    lock = threading.Lock()

    def work():
        while True:
            with lock:
                if not jobs:
                    return
                f = jobs.pop()
            lib.oracle_synth_rows(C.c_void_p(out[f].ctypes.data), 0, height, width, kind, seed, first_frame + f)

    ts = [threading.Thread(target=work) for _ in range(max(1, min(threads, n_frames)))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return out


class Ref(_Base):
    prefix = "ref_"

    def __init__(self, path: Path = REF_LIB):
        super().__init__(path)
        self.lib.ref_canny.restype = C.c_double

    @staticmethod
    def available() -> bool:
        return REF_LIB.exists()

    def canny(self, img, sigma: float, lo: int, hi: int, want_edges: bool = True):
        a = _c(img, np.uint8)
        h, w = a.shape
        edges = np.empty((h, w), np.int16) if want_edges else None
        self.last_seconds = self.lib.ref_canny(_p(a), C.c_float(sigma), int(lo), int(hi), h, w,
                                               _p(edges) if want_edges else None)
        return edges
