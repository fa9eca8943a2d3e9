"""The error bound behind pcf2d's float32 rotation (amep_b200/csrc/pcf2d.cu, PairBinner::bins).

The kernel evaluates a rotated difference vector in float32 first and trusts the result only if it
is farther than `1.01e-6 (|ux| + |uy| + |cx| + |cy|) + 2.5e-7 max|edge|` from a bin threshold.  The
claim it rests on: the float32 value differs from the float64 value the reference bins
(amep/utils.py:581-652 applied to float32 differences) by at most
`4e-7 (|ux| + |uy|) + 2.5e-7 (|cx| + |cy|)`.  Checked here in NumPy, with the kernel's operation
order (one fused multiply-add, emulated exactly in float64) and with separately rounded
operations, for float32 and float64 centres, over magnitudes from 1e-3 to 1e6."""
import zlib

import numpy as np
import pytest

f32 = np.float32


def _float32_rotation(dx, dy, cx, cy, c, s, fused):
    """xa, ya as the kernel computes them: u = d - fl32(centre); fma(r00, ux, r01*uy) + fl32(cx)."""
    fcx, fcy = f32(cx), f32(cy)
    r00, r01, r10, r11 = f32(c), f32(-s), f32(s), f32(c)
    ux, uy = (dx - fcx).astype(f32), (dy - fcy).astype(f32)
    p01, p11 = (r01 * uy).astype(f32), (r11 * uy).astype(f32)
    if fused:   # the float64 product of two float32 values is exact, so this is a true fma
        tx = (r00.astype(np.float64) * ux.astype(np.float64) + p01.astype(np.float64)).astype(f32)
        ty = (r10.astype(np.float64) * ux.astype(np.float64) + p11.astype(np.float64)).astype(f32)
    else:
        tx = ((r00 * ux).astype(f32) + p01).astype(f32)
        ty = ((r10 * ux).astype(f32) + p11).astype(f32)
    return (tx + fcx).astype(f32), (ty + fcy).astype(f32), ux, uy


def _float64_rotation(dx, dy, cx, cy, c, s, centre_is_f32):
    """The reference: v = diff - centre (float32 subtraction if both are float32), R.v and + centre
    in float64 with separately rounded products and sums."""
    if centre_is_f32:
        vx, vy = (dx - f32(cx)).astype(f32).astype(np.float64), (dy - f32(cy)).astype(f32).astype(np.float64)
        cx, cy = float(f32(cx)), float(f32(cy))
    else:
        vx, vy = dx.astype(np.float64) - cx, dy.astype(np.float64) - cy
    return (c * vx + (-s) * vy) + cx, (s * vx + c * vy) + cy


@pytest.mark.parametrize("centre_is_f32", [True, False])
@pytest.mark.parametrize("fused", [True, False])
def test_float32_rotation_error_bound(centre_is_f32, fused):
    rng = np.random.default_rng(zlib.crc32(f"margin-{centre_is_f32}-{fused}".encode()))
    worst = 0.0
    for scale in (1e-3, 1.0, 37.0, 1e3, 1e6):
        for cscale in (0.0, 1.0, 512.3, 1e5):
            n = 200_000
            dx = (rng.uniform(-1, 1, n) * scale).astype(f32)
            dy = (rng.uniform(-1, 1, n) * scale).astype(f32)
            cx, cy = rng.uniform(-1, 1, 2) * cscale
            ang = rng.uniform(0, 2 * np.pi)
            c, s = np.cos(-ang), np.sin(-ang)
            xa, ya, ux, uy = _float32_rotation(dx, dy, cx, cy, c, s, fused)
            x, y = _float64_rotation(dx, dy, cx, cy, c, s, centre_is_f32)
            bound = 4e-7 * (np.abs(ux.astype(np.float64)) + np.abs(uy.astype(np.float64))) + 2.5e-7 * (abs(cx) + abs(cy))
            err = np.maximum(np.abs(xa.astype(np.float64) - x), np.abs(ya.astype(np.float64) - y))
            ok = err <= bound + 1e-45
            assert np.all(ok), (scale, cscale, float(np.max(err[~ok] / bound[~ok])))
            with np.errstate(divide="ignore", invalid="ignore"):
                ratio = np.where(bound > 0, err / bound, 0.0)
            worst = max(worst, float(np.max(ratio)))
    # the bound is not vacuous either: the observed error comes within a small factor of it
    assert 0.05 < worst <= 1.0, worst
