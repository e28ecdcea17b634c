"""Sizing study of BASELINE's network configs with the greedy path (no GPU needed)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from tnc_b200.builders import random_circuit, sycamore_circuit
from tnc_b200.contractionpath.paths import Cotengrust
from tnc_b200.tensornetwork import Tensor


def analyse(tn, name):
    t0 = time.time()
    opt = Cotengrust(tn); opt.find_path()
    p = opt.get_best_replace_path()
    ts = list(tn.tensors)
    tot = 0.0; big = []; small = 0; peak = 0.0; live = sum(t.size() for t in ts)
    maxrank = 0
    for (i, j) in p.toplevel:
        a, b = ts[i], ts[j]
        k = (a & b).size(); o = a ^ b
        m = (a - b).size(); n = (b - a).size()
        fl = 8 * m * n * k; tot += fl
        if m * n * k <= 1024: small += 1
        big.append((fl, int(np.log2(m)), int(np.log2(n)), int(np.log2(k)), len(a.legs), len(b.legs), len(o.legs)))
        live += o.size(); peak = max(peak, live); live -= a.size() + b.size()
        maxrank = max(maxrank, len(o.legs))
        ts[i] = o; ts[j] = Tensor()
    big.sort(reverse=True)
    print(f"{name}: leaves {len(tn.tensors)} pairs {len(p.toplevel)} flops {tot:.3e} peak {peak*16/2**30:.2f} GiB maxrank {maxrank} small(<=2^10) {small} pathfind {time.time()-t0:.1f}s")
    for b in big[:4]:
        print(f"     {b[0]:.2e} flop ({100*b[0]/tot:.0f}%) M=2^{b[1]} N=2^{b[2]} K=2^{b[3]} ranks {b[4]}x{b[5]}->{b[6]}")


if __name__ == "__main__":
    for q, r, seed in [(24, 10, 1), (24, 12, 1), (24, 14, 1), (30, 12, 1), (36, 10, 1), (36, 12, 1), (36, 14, 1)]:
        tn = random_circuit(q, r, 0.5, 0.5, np.random.default_rng(seed))
        analyse(tn, f"random {q}q {r}r seed{seed}")
    for d in (6, 8, 10):
        c = sycamore_circuit(53, d, np.random.default_rng(1))
        analyse(c.into_amplitude_network("0" * 53)[0], f"sycamore53 depth{d}")
