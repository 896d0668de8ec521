"""Integer model of the tcgen05 (int8-sliced) fp64 update of egobox_b200/csrc/kernels_ozaki.cu  --  TEST INFRASTRUCTURE ONLY.

Not a restatement of reference code (the reference multiplies in fp64, gp/src/algorithm.rs:1004 / :1077): this is the CPU
model of OUR kernel's arithmetic, digit for digit, so that its error bound is pinned without a GPU:
  scale     s_i = 2^(ilogb(max|row i|) + 3)            (ozaki_rowscale_kernel / the quarter-row maxima of the panel solves)
  integer   t   = rint(x * 2^56 / s_i),  |t| < 2^54
  digits    the 7 low bytes of (t + 0x0080808080808080), each XOR 0x80: balanced base-256 digits d_6 .. d_0
  products  ACC_w = sum_{p+q=w} S_p S_q^T  for w = 0..6 (exact in int32: < 2^25)
  fold      pass 0: T0 = ((a0*256 + a1)*256 + a2)*256 + a3, scale 256^-5 ; pass 1: T1 = (a4*256 + a5)*256 + a6, scale 256^-8
  result    (P P^T)_ij ~= s_i s_j (T0 256^-5 + T1 256^-8)
Only tests/ may import it."""
from __future__ import annotations

import numpy as np

BIAS = 0x0080808080808080
N_DIGITS = 7


def row_scales(p):
    m = np.abs(p).max(axis=1)
    s = np.zeros_like(m)
    nz = (m > 0) & (m < 1e300)
    s[nz] = np.ldexp(1.0, np.frexp(m[nz])[1] - 1 + 3)       # ilogb(m) = frexp exponent - 1
    return s


def digits(p, s):
    """(7, rows, K) int8 digit slices, most significant first, and the integers t they encode."""
    inv = np.where(s > 0, np.ldexp(1.0, 56) / np.where(s > 0, s, 1.0), 0.0)
    t = np.rint(p * inv[:, None]).astype(np.int64)
    u = (t + np.int64(BIAS)).astype(np.int64)
    d = np.empty((N_DIGITS,) + p.shape, dtype=np.int8)
    for byte in range(N_DIGITS):                              # byte j = digit 6 - j
        b = ((u >> np.int64(8 * byte)) & np.int64(255)).astype(np.uint8) ^ np.uint8(0x80)
        d[N_DIGITS - 1 - byte] = b.view(np.int8)
    return d, t


def sliced_product(a, b):
    """fp64 model of C = A B^T through the digit slices; returns (C, dict of checks)."""
    sa, sb = row_scales(a), row_scales(b)
    da, ta = digits(a, sa)
    db, tb = digits(b, sb)
    acc = []
    for w in range(N_DIGITS):
        m = np.zeros((a.shape[0], b.shape[0]), dtype=np.int64)
        for p in range(w + 1):
            q = w - p
            m += da[p].astype(np.int64) @ db[q].astype(np.int64).T
        acc.append(m)
    t0 = ((acc[0] * 256 + acc[1]) * 256 + acc[2]) * 256 + acc[3]
    t1 = (acc[4] * 256 + acc[5]) * 256 + acc[6]
    sij = sa[:, None] * sb[None, :]
    c = t0.astype(np.float64) * (sij * 256.0 ** -5) + t1.astype(np.float64) * (sij * 256.0 ** -8)
    return c, dict(acc_max=max(int(np.abs(m).max()) for m in acc), t0_max=int(np.abs(t0).max()), ta=ta, tb=tb, da=da, db=db)
