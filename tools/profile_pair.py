"""One pair through the default engine a few times -- the target of `ncu -k regex:crt_gemm ...` captures.
usage: python tools/profile_pair.py MxNxK [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tnc_b200 as tb

M, N, K = (int(x) for x in sys.argv[1].split("x"))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = tb.Context(0)
rng = np.random.default_rng(0)
a = tb.DeviceTensor.from_numpy(ctx, rng.standard_normal((K, M)) + 1j * rng.standard_normal((K, M)))
b = tb.DeviceTensor.from_numpy(ctx, rng.standard_normal((N, K)) + 1j * rng.standard_normal((N, K)))
c = tb.DeviceTensor.empty(ctx, (N, M))
for _ in range(reps):
    tb.contract_pair_into(ctx, [0, 1], a, [2, 0], b, c)
ctx.synchronize()
print(ctx.engine_counts(), ctx.last_tcgen05_info())
