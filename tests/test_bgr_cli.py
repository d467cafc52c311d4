"""The two pieces either side of the hot path that SURVEY 8(f) marks "next": BGR->gray on the device (the reference does
cvtColor on the host, src/main.cpp:113) and a file front end with the reference CLI's arguments (src/main.cpp:29-76)."""
import numpy as np
import pytest

import canny_edge_b200 as cb
from canny_edge_b200 import cli


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
