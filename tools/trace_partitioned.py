"""Per-step device times (TNCB_TRACE) of the committed N-part plan of the bench network executed on ONE GPU.
usage: TNCB_TRACE=1 python tools/trace_partitioned.py N"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import plan_partitions as pp
import tnc_b200 as tb
from tnc_b200.tensornetwork import contract_tensor_network

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tn = pp.build_network()
ptn, path, facts = pp.load(tn, n)
ctx = tb.Context(0)
for _ in range(2):
    t0 = time.perf_counter()
    amp = complex(contract_tensor_network(ptn, path, ctx=ctx).to_numpy())
    print("partitioned path on one GPU: %.2f ms, amplitude %s" % ((time.perf_counter() - t0) * 1e3, amp), flush=True)
