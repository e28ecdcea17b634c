"""Per-step device times (TNCB_TRACE) of ONE slice of a sliced path file (tools/search_path.py) on one GPU.
usage: TNCB_TRACE=1 python tools/trace_slice.py bench_inputs/sycamore53_d12.json 2> trace.txt"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_sliced import build
import tnc_b200 as tb
from tnc_b200.contractionpath import ContractionPath
from tnc_b200.contractionpath.slicing import SlicedNetwork
from tnc_b200.tensornetwork import contract_tensor_network

d = json.load(open(sys.argv[1]))
tn = build(d["network"])
path = ContractionPath.simple([tuple(x) for x in d["toplevel"]])
sn = SlicedNetwork(tn, d["sliced_legs"])
ctx = tb.Context(0)
for rep in range(2):
    t0 = time.perf_counter()
    amp = complex(contract_tensor_network(sn.slice(sn.assignments[0]), path, ctx=ctx).to_numpy())
    print("slice 0: %.2f ms, partial amplitude %s" % ((time.perf_counter() - t0) * 1e3, amp), flush=True)
    print("TNCB_TRACE ---- end of repetition %d" % rep, file=sys.stderr, flush=True)
