"""The C++ drop-in header (include/canny_b200_compat.hpp): compiles against the C ABI with plain g++ (CPU
check) and, on a B200, reproduces the oracle through the reference's own calling convention."""
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

import canny_edge_b200 as cb

ROOT = Path(__file__).resolve().parent.parent


def build_exe(tmp: Path, name: str = "compat_main") -> Path:
    exe = tmp / name
    lib_dir = ROOT / "canny_edge_b200"
    cmd = ["g++", "-std=c++14", "-O2", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "cpp" / f"{name}.cpp"), "-o", str(exe),
           f"-L{lib_dir}", "-lcanny_b200", f"-Wl,-rpath,{lib_dir}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_compat_header_compiles_and_fails_loudly_without_gpu(tmp_path):
    cb.load()
    exe = build_exe(tmp_path)
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the gpu test covers the run")
    img = np.zeros((8, 8), np.uint8)
    img.tofile(tmp_path / "in.u8")
    r = subprocess.run([str(exe), str(tmp_path / "in.u8"), "8", "8", "1.4", "20", "60", str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr  # no silent CPU fallback


@pytest.mark.gpu
def test_compat_header_matches_oracle(tmp_path, oracle):
    exe = build_exe(tmp_path)
    h, w = 203, 318
    img = cb.synth_host(1, h, w, kind=0, seed=99)[0]
    img.tofile(tmp_path / "in.u8")
    r = subprocess.run([str(exe), str(tmp_path / "in.u8"), str(h), str(w), "1.4", "20", "60", str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = dict(zip(("blur", "mag", "ang", "nms", "edges"), oracle.canny(img, 1.4, 20, 60, steps=True)))
    for name in ("blur", "mag", "ang", "nms", "edges"):
        got = np.fromfile(str(tmp_path / f"o.{name}.i16"), np.int16).reshape(h, w)
        assert (got == want[name]).all(), name
    got = np.fromfile(str(tmp_path / "o.edges2.i16"), np.int16).reshape(h, w)
    assert (got == want["edges"]).all()


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_bands_cpp_caller_compiles_and_fails_loudly_without_gpu(tmp_path):
    cb.load()
    exe = build_exe(tmp_path, "bands_main")
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the gpu test covers the run")
    np.zeros((64, 64), np.uint8).tofile(tmp_path / "in.u8")
    r = subprocess.run([str(exe), str(tmp_path / "in.u8"), "64", "64", "1.4", "20", "60", "2", "1", str(tmp_path / "o.u8")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_bands_cpp_caller_matches_oracle(tmp_path, oracle):
    """The multi-GPU band handle driven from plain C++ (tests/cpp/bands_main.cpp): in-process group of 5 bands, two steps."""
    exe = build_exe(tmp_path, "bands_main")
    h, w = 700, 640
    img = cb.synth_host(1, h, w, kind=1, seed=5)[0]
    img.tofile(tmp_path / "in.u8")
    r = subprocess.run([str(exe), str(tmp_path / "in.u8"), str(h), str(w), "1.4", "20", "60", "5", "2", str(tmp_path / "o.u8")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "transport=1" in r.stdout
    got = np.fromfile(str(tmp_path / "o.u8"), np.uint8).reshape(h, w)
    assert (got.astype(np.int16) == oracle.canny(img, 1.4, 20, 60)).all()
