"""HBM bandwidth of tncb_permute (K3 tiled transpose vs the plain gather kernel, TNCB_NO_K3=1) on statevector-like
and matrix-like permutations: 32 bytes of traffic per element.  usage: python tools/bench_permute.py"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import tnc_b200 as tb

ctx = tb.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=0)
cases = [("matrix 8192x8192 transpose", (8192, 8192), (1, 0)),
         ("26 qubit legs reversed", (2,) * 26, tuple(reversed(range(26)))),
         ("26 qubit legs, Permutor-like (pairs swapped)", (2,) * 26, tuple(i ^ 1 for i in range(26))),
         ("rank-12 dim-4 interleave (C2 operand to GEMM order)", (4,) * 12, (0, 2, 4, 6, 8, 10, 11, 9, 7, 5, 3, 1)),
         ("3D 512x384x256 -> (2,0,1)", (512, 384, 256), (2, 0, 1))]
for name, shape, perm in cases:
    n = int(np.prod(shape))
    d = tb.DeviceTensor.empty(ctx, shape)
    times = []
    for rep in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        out = C.c_void_p()
        e0.record(stream)
        tb.check(ctx._l.tncb_permute(ctx.handle, d.handle, (C.c_int * len(perm))(*perm), C.byref(out)))
        e1.record(stream); ctx.synchronize(); torch.cuda.synchronize()
        d.release()
        d = tb.DeviceTensor.adopt(ctx, out)
        d.shape = tuple(shape)          # keep permuting a tensor of the same shape (the content does not matter)
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times[1:]))
    print(json.dumps({"case": name, "elements": n, "ms": round(ms, 4), "GBps": round(32.0 * n / ms * 1e-6, 1),
                      "engine": "plain gather" if os.environ.get("TNCB_NO_K3") else "K3 tiled"}), flush=True)
    d.free()
