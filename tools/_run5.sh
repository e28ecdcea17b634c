set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; tail -3 gpurun_out/r02_bench_c.err; cut -c1-700 gpurun_out/r02_bench_c.json
for n in 2 4 8; do TNCB_TRACE=1 timeout 300 python tools/trace_partitioned.py $n 2> gpurun_out/r02_trace_part$n.txt; done
python - <<'PY'
import re
for n in (2,4,8):
    rows=[]
    for l in open(f'gpurun_out/r02_trace_part{n}.txt'):
        m=re.search(r'step (\d+) class K(\d) M (\d+) N (\d+) K (\d+).* ms ([0-9.]+) tflops ([0-9.]+)',l)
        if m: rows.append((int(m.group(2)),int(m.group(3)),int(m.group(4)),int(m.group(5)),float(m.group(6)),float(m.group(7))))
    half=rows[len(rows)//2:]
    print(n,'steps',len(half),'sum ms',sum(r[4] for r in half))
    for r in sorted(half,key=lambda r:-r[4])[:10]: print('   ',r)
PY
timeout 300 python tools/bench_network.py --qubits 20 --rounds 8 --seed 4 --steps 20 --plan 2>&1 | cut -c1-500
timeout 300 python tools/bench_network.py --qubits 20 --rounds 8 --seed 4 --steps 20 2>&1 | cut -c1-500
TNCB_NO_BATCH=1 timeout 300 python tools/bench_network.py --qubits 20 --rounds 8 --seed 4 --steps 20 --plan 2>&1 | cut -c1-500
