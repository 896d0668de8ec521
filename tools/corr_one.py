"""One correlation build for an ncu capture:  python tools/corr_one.py <corr 0..3> <d> [n]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import egobox_b200 as eg                                   # noqa: E402
from tools._util import make_problem, make_context         # noqa: E402

corr, d = int(sys.argv[1]), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
x, y = make_problem(n, d, seed=42)
ctx = make_context(x, y, corr, eg.CONSTANT)
for _ in range(2):
    print(ctx.reduced_likelihood(np.full(d, 1.0)))
ctx.close()
