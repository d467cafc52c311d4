"""The two pieces either side of the hot path that SURVEY 8(f) marks "next": BGR->gray on the device (the reference does
cvtColor on the host, src/main.cpp:113) and a file front end with the reference CLI's arguments (src/main.cpp:29-76)."""
import ctypes as C

import numpy as np
import pytest

import canny_edge_b200 as cb
from canny_edge_b200 import cli
from canny_edge_b200._lib import check


def test_cli_argument_validation_mirrors_main_cpp():
    ok = cli.parse_args(["1.4", "20", "60", "a.png", "-s", "b.png", "-c"])
    assert ok[:4] == (1.4, 20, 60, True) and ok[5] == ["a.png", "b.png"]
    for argv, msg in ((["1.4", "60", "20", "x.png"], "minVal must be less than maxVal"),      # src/main.cpp:63-66
                      (["1.4", "-1", "20", "x.png"], "minVal must be in the range"),          # :68-71
                      (["1.4", "20", "256", "x.png"], "maxVal must be in the range"),         # :73-76
                      (["1.4", "20", "60"], "USAGE"), (["1.4", "20"], "USAGE")):              # :47-56
        with pytest.raises(cli.CliError) as e:
            cli.parse_args(argv)
        assert msg in str(e.value)


def gray_formula(bgr):
    b, g, r = (bgr[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def test_gray_formula_is_opencvs():
    cv2 = pytest.importorskip("cv2")
    t = np.random.default_rng(0).integers(0, 256, (600, 500, 3)).astype(np.uint8)
    assert (gray_formula(t) == cv2.cvtColor(t, cv2.COLOR_BGR2GRAY)).all()


@pytest.mark.gpu
def test_canny_bgr_matches_cvtcolor_plus_oracle(gpu_ctx, oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for h, w in ((64, 64), (241, 322), (480, 641), (3301, 3403)):   # the last one (>= 32 MB of BGR) takes the staged / bit-packed host path
        gray0 = cb.synth_host(1, h, w, kind=0, seed=h)[0].astype(np.int16)
        frame = np.stack([np.clip(gray0 + rng.integers(-40, 41, (h, w)), 0, 255) for _ in range(3)], axis=-1).astype(np.uint8)
        edges, gray = cb.cuda_canny_bgr(frame, 1.4, 20, 60, return_gray=True, ctx=gpu_ctx)
        want_gray = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)
        assert (gray == want_gray).all() and (gray == gray_formula(frame)).all()
        assert (edges == oracle.canny(want_gray, 1.4, 20, 60)).all()


def _colour_frames(n, h, w, seed):
    rng = np.random.default_rng(seed)
    gray0 = cb.synth_host(n, h, w, kind=0, seed=seed).astype(np.int16)
    return np.stack([np.clip(gray0 + rng.integers(-40, 41, (n, h, w)), 0, 255) for _ in range(3)], axis=-1).astype(np.uint8)


@pytest.mark.gpu
@pytest.mark.parametrize("sigma", [0.5, 1.4, 2.0, 3.0])
def test_fused_bgr_front_kernel_matches_cvtcolor_plus_oracle(gpu_ctx, oracle, sigma, monkeypatch):
    """b200_canny_batch_device_bgr: BGR -> gray inside the front kernel's staging.  Shapes that take the fused
    kernel (width % 16 == 0; half-windows 2, 5, 6, 9), shapes that cannot (ragged width), strips at both image borders, bands of
    several slabs, more frames than one chunk; every map against cvtColor + the oracle."""
    import torch
    cv2 = pytest.importorskip("cv2")
    for n, h, w in ((1, 64, 64), (3, 200, 272), (2, 333, 1040), (1, 1080, 1920), (2, 97, 250), (11, 70, 128)):
        frames = _colour_frames(n, h, w, seed=h + n)
        d_in = torch.from_numpy(frames).cuda()
        d_out = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        gpu_ctx.set_chunk_frames(4 if n > 4 else 0)
        fast0, gen0 = gpu_ctx.front_kernel_stats()
        cb.canny_batch_device_bgr_ptr(gpu_ctx, d_in.data_ptr(), n, h, w, sigma, 20, 60, d_out.data_ptr())
        gpu_ctx.synchronize()
        gpu_ctx.set_chunk_frames(0)
        got = d_out.cpu().numpy()
        for f in range(n):
            gray = cv2.cvtColor(frames[f], cv2.COLOR_BGR2GRAY)
            assert (gray == gray_formula(frames[f])).all()
            want = oracle.canny(gray, sigma, 20, 60)
            assert (got[f].astype(np.int16) == want).all(), f"{n}x{h}x{w} sigma {sigma} frame {f}: {(got[f] != want).sum()} px differ"
        assert gpu_ctx.front_kernel_stats()[1] == gen0           # never the generic kernel


@pytest.mark.gpu
def test_canny_bgr_without_gray_plane_takes_the_fused_kernel(gpu_ctx, oracle):
    """b200_canny_bgr with gray_out == NULL: host B,G,R frame in, int16 map out, conversion inside the front kernel where the width
    allows (640, 1920) and through the conversion kernel where it does not (641)."""
    cv2 = pytest.importorskip("cv2")
    for h, w in ((480, 640), (480, 641), (1080, 1920)):
        frame = _colour_frames(1, h, w, seed=w)[0]
        edges = cb.cuda_canny_bgr(frame, 1.4, 20, 60, ctx=gpu_ctx)
        want = oracle.canny(cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY), 1.4, 20, 60)
        assert (edges == want).all(), f"{h}x{w}: {(edges != want).sum()} px differ"


@pytest.mark.gpu
def test_fused_bgr_equals_separate_conversion(gpu_ctx, monkeypatch):
    """Full 4K frames: the fused kernel and (conversion pass + gray kernel) give the same bytes."""
    import torch
    n, h, w = 3, 2160, 3840
    frames = _colour_frames(n, h, w, seed=5)
    d_in = torch.from_numpy(frames).cuda()
    d_fused = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    d_gray = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    d_sep = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    cb.canny_batch_device_bgr_ptr(gpu_ctx, d_in.data_ptr(), n, h, w, 1.4, 20, 60, d_fused.data_ptr())
    check(cb.load().b200_bgr_to_gray_device(gpu_ctx.handle, C.c_void_p(d_in.data_ptr()), n * h * w, C.c_void_p(d_gray.data_ptr())))
    cb.canny_batch_device_ptr(gpu_ctx, d_gray.data_ptr(), n, h, w, 1.4, 20, 60, d_sep.data_ptr())
    gpu_ctx.synchronize()
    assert torch.equal(d_fused, d_sep)
    assert int((d_fused == 255).sum()) > 10000


@pytest.mark.gpu
def test_cli_end_to_end(tmp_path, oracle, test_gray):
    cv2 = pytest.importorskip("cv2")
    src = tmp_path / "test.png"
    cv2.imwrite(str(src), np.stack([test_gray] * 3, axis=-1))       # gray image stored as colour: BGR2GRAY gives it back exactly
    assert cli.run(["1.4", "20", "60", "-s", "-o", str(tmp_path / "out"), str(src)]) == 0
    got = cv2.imread(str(tmp_path / "out" / "test_edges.png"), cv2.IMREAD_GRAYSCALE)
    assert (got.astype(np.int16) == oracle.canny(test_gray, 1.4, 20, 60)).all()     # 3466 edge pixels, SURVEY 8(c)
    assert int((got == 255).sum()) == 3466
    for name in ("blur", "magnitude", "nms"):
        assert (tmp_path / "out" / f"test_{name}.png").exists()
