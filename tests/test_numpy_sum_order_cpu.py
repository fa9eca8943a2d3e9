"""The two float32 summation orders of NumPy that amepb_sq_debye_ref reproduces on the GPU
(amep/spatialcor.py:1385-1397): ``np.sum(a)`` of a contiguous 1D float32 array is the pairwise
summation of numpy/_core/src/umath/loops_utils.h.src (leaves of <= 128 elements with eight strided
accumulators, ranges split at n/2 - (n/2) % 8), ``np.sum(a, axis=0)`` of a C-contiguous 2D array adds
the rows one after the other."""
import numpy as np
import pytest

f32 = np.float32


def pairwise(a):
    n = len(a)
    if n < 8:
        r = f32(0)
        for x in a:
            r = f32(r + x)
        return r
    if n <= 128:
        r = [a[j] for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = f32(r[j] + a[i + j])
            i += 8
        res = f32(f32(f32(r[0] + r[1]) + f32(r[2] + r[3])) + f32(f32(r[4] + r[5]) + f32(r[6] + r[7])))
        while i < n:
            res = f32(res + a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return f32(pairwise(a[:n2]) + pairwise(a[n2:]))


@pytest.mark.parametrize("n", [1, 5, 8, 9, 100, 128, 129, 257, 1000, 4097, 100003])
def test_np_sum_float32_is_pairwise(n):
    a = np.random.default_rng(n).standard_normal(n).astype(f32)
    assert np.sum(a) == pairwise(a)


def test_np_sum_axis0_adds_rows_in_order():
    a = np.random.default_rng(0).standard_normal((5000, 7)).astype(f32)
    acc = a[0].copy()
    for i in range(1, len(a)):
        acc = (acc + a[i]).astype(f32)
    assert np.array_equal(np.sum(a, axis=0), acc)
