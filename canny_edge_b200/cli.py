"""File front end with the reference CLI's arguments.

The reference's `main` (src/main.cpp:18-142) takes `sigma minVal maxVal [-c] [-s]`, grabs one webcam frame, converts it to gray
(src/main.cpp:113) and shows the Canny result; a TODO there (src/main.cpp:108-110) asks for more than one frame.  This front end
keeps the three positional parameters, their validation and the `-s` (steps) switch, and reads image FILES instead (host-side
OpenCV for decode / encode only — the arithmetic, including BGR->gray, runs in libcanny_b200.so):

    python -m canny_edge_b200.cli 1.4 20 60 [-s] [-o OUTDIR] image [image ...]

For every input it writes OUTDIR/<stem>_edges.png (and with -s also <stem>_blur.png, <stem>_magnitude.png, <stem>_nms.png, the
planes the reference displays after each stage, min-max stretched to 8 bit as src/cuda.cu:404-405 does).  `-c` is accepted and
ignored: there is no CPU path here.
"""
from __future__ import annotations

import sys
from pathlib import Path
from typing import List, Optional, Tuple

import numpy as np

USAGE = ("USAGE: {prog} sigma minVal maxVal [-s] [-o OUTDIR] image [image ...]\n"
         "   sigma: Standard deviation used for the gaussian blurring kernel\n"
         "   minVal: The minimum threshold value used for hysteresis\n"
         "           Must be in the range of [0,255]\n"
         "   maxVal: The maximum threshold value used for hysteresis\n"
         "           Must be in the range of [0,255]\n")


class CliError(Exception):
    pass


def parse_args(argv: List[str]) -> Tuple[float, int, int, bool, Path, List[str]]:
    """Mirrors src/main.cpp:29-76: flags anywhere, three positional values, the same range checks and messages."""
    steps, out_dir, values = False, Path("."), []
    it = iter(argv)
    for arg in it:
        if arg == "-c":
            continue                      # the reference's "use CUDA" switch: always on here
        elif arg == "-s":
            steps = True
        elif arg == "-o":
            try:
                out_dir = Path(next(it))
            except StopIteration:
                raise CliError("ERROR: -o needs a directory") from None
        else:
            values.append(arg)
    if len(values) < 4:
        raise CliError(USAGE.format(prog="canny_edge_b200.cli"))
    try:
        sigma, lo, hi = float(values[0]), int(values[1]), int(values[2])
    except ValueError:
        raise CliError(USAGE.format(prog="canny_edge_b200.cli")) from None
    if hi <= lo:
        raise CliError("ERROR: minVal must be less than maxVal")
    if lo < 0 or lo > 255:
        raise CliError("ERROR: minVal must be in the range of [0,255]")
    if hi < 0 or hi > 255:
        raise CliError("ERROR: maxVal must be in the range of [0,255]")
    if not sigma > 0:
        raise CliError("ERROR: sigma must be positive")
    return sigma, lo, hi, steps, out_dir, values[3:]


def _stretch(plane: np.ndarray) -> np.ndarray:
    """cv::normalize(..., 0, 255, NORM_MINMAX) + convertTo(CV_8U), the reference's display scaling (src/cuda.cu:404-405)."""
    p = plane.astype(np.float64)
    lo, hi = float(p.min()), float(p.max())
    if hi <= lo:
        return np.zeros(plane.shape, np.uint8)
    return np.clip(np.rint((p - lo) * (255.0 / (hi - lo))), 0, 255).astype(np.uint8)


def run(argv: Optional[List[str]] = None) -> int:
    import cv2

    from . import api

    try:
        sigma, lo, hi, steps, out_dir, files = parse_args(list(sys.argv[1:] if argv is None else argv))
    except CliError as e:
        sys.stderr.write(str(e) + "\n")
        return 0                           # the reference exits with status 0 on usage errors (src/main.cpp:37,56,66)
    out_dir.mkdir(parents=True, exist_ok=True)
    ctx = api.Context(0)
    try:
        for name in files:
            frame = cv2.imread(name, cv2.IMREAD_COLOR)
            if frame is None:
                sys.stderr.write(f"ERROR: cannot read {name}\n")
                continue
            stem = Path(name).stem
            edges, gray = api.cuda_canny_bgr(frame, sigma, lo, hi, return_gray=True, ctx=ctx)
            cv2.imwrite(str(out_dir / f"{stem}_edges.png"), edges.astype(np.uint8))
            if steps:
                blur, mag, _, nms, _ = api.cuda_canny(gray, sigma, lo, hi, steps=True, ctx=ctx)
                cv2.imwrite(str(out_dir / f"{stem}_blur.png"), _stretch(blur))
                cv2.imwrite(str(out_dir / f"{stem}_magnitude.png"), _stretch(mag))
                cv2.imwrite(str(out_dir / f"{stem}_nms.png"), _stretch(nms))
            print(f"{name}: {frame.shape[1]}x{frame.shape[0]}, {int((edges == 255).sum())} edge pixels")
    finally:
        ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(run())
