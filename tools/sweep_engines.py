"""DMMA (K1) vs tcgen05 int8 slicing (K1') over GEMM shapes: device time per pair, same box."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tnc_b200 as tb

ctx = tb.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=0)
shapes = [(512,512,512),(1024,1024,1024),(1536,1536,1536),(2048,2048,2048),(1024,1024,4096),(4096,4096,512),(4096,4096,256),
          (2048,512,2048),(512,4096,4096),(8192,8192,1024),(16384,8192,4096)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in s.split("x")) for s in sys.argv[1:]]
rng = np.random.default_rng(0)
for (M, N, K) in shapes:
    a = tb.DeviceTensor.from_numpy(ctx, (rng.standard_normal((K, M)) + 1j * rng.standard_normal((K, M))))
    b = tb.DeviceTensor.from_numpy(ctx, (rng.standard_normal((N, K)) + 1j * rng.standard_normal((N, K))))
    c = tb.DeviceTensor.empty(ctx, (N, M))
    out = {"M": M, "N": N, "K": K}
    ref = None
    for s in (0, 8, 7, 6):
        ctx.set_tcgen05_slices(s)
        os.environ["TNCB_FORCE_TCGEN05"] = "1"
        for _ in range(2):
            tb.contract_pair_into(ctx, [0, 1], a, [2, 0], b, c)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record(stream)
        for _ in range(reps):
            tb.contract_pair_into(ctx, [0, 1], a, [2, 0], b, c)
        e1.record(stream); ctx.synchronize(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res = c.to_numpy() if M * N <= 2048 * 2048 else None
        if s == 0:
            ref = res
            out["dmma_ms"] = round(ms, 4); out["dmma_tf"] = round(8.0 * M * N * K / ms * 1e-9, 1)
        else:
            out[f"s{s}_ms"] = round(ms, 4); out[f"s{s}_tf"] = round(8.0 * M * N * K / ms * 1e-9, 1)
            if res is not None:
                out[f"s{s}_err"] = float(np.abs(res - ref).max() / np.abs(ref).max())
    print(json.dumps(out), flush=True)
    a.free(); b.free(); c.free()
