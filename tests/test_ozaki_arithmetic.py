"""The integer arithmetic of the tcgen05 (int8-sliced) fp64 update, modelled on the CPU (oracle/ozaki_oracle.py): digit
reconstruction is exact, the partial products fit int32, the folded integers fit the 2^51 range of the conversion trick, and
the result is as close to the exact product as a plain fp64 dot product."""
from fractions import Fraction

import numpy as np

from oracle import ozaki_oracle as Z


def _panel(rows, k, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1, 1, (rows, k)) * 10.0 ** -rng.uniform(0, 6, (rows, k))


def test_digits_reconstruct_the_integer_exactly_and_fill_int8():
    p = _panel(64, 256, 1)
    p[5] = 0.0                                   # an all-zero (padding) row
    s = Z.row_scales(p)
    d, t = Z.digits(p, s)
    assert s[5] == 0.0 and not d[:, 5].any()
    rec = np.zeros_like(t)
    for k in range(Z.N_DIGITS):
        rec = rec * 256 + d[k].astype(np.int64)
    np.testing.assert_array_equal(rec, t)
    assert d.min() >= -128 and d.max() <= 127
    assert np.abs(t).max() < 2 ** 54             # 4 |row|max <= s
    assert np.abs(d[0]).max() <= 64              # the leading digit never overflows
    # x = s * t / 2^56 up to the rounding of t (half a unit of 2^-56 s)
    err = np.abs(p - s[:, None] * t / 2.0 ** 56)
    assert np.all(err <= s[:, None] * 2.0 ** -57 * (1 + 1e-12))


def test_partial_products_fit_int32_and_fold_fits_the_conversion_range():
    a, b = _panel(128, 256, 2), _panel(128, 256, 3)
    a[:, :] = np.where(np.abs(a) > 0, np.sign(a), 1.0) * np.abs(a).max()     # worst case: every entry at the row maximum
    _, chk = Z.sliced_product(a, b)
    assert chk["acc_max"] < 2 ** 31
    assert chk["t0_max"] < 2 ** 51


def test_sliced_product_is_as_accurate_as_an_fp64_dot_product():
    a, b = _panel(48, 256, 4), _panel(40, 256, 5)
    c, _ = Z.sliced_product(a, b)
    plain = a @ b.T
    worst_sliced = worst_plain = 0.0
    for i in range(0, 48, 5):
        for j in range(0, 40, 7):
            exact = sum(Fraction(float(x)) * Fraction(float(y)) for x, y in zip(a[i], b[j]))
            den = 256 * np.abs(a[i]).max() * np.abs(b[j]).max()
            worst_sliced = max(worst_sliced, abs(float(Fraction(float(c[i, j])) - exact)) / den)
            worst_plain = max(worst_plain, abs(float(Fraction(float(plain[i, j])) - exact)) / den)
    assert worst_sliced < 2.0 ** -50             # far inside the 2^-47 s_i s_j bound of the dropped digit pairs
    assert worst_sliced < 4 * max(worst_plain, 2.0 ** -56)
