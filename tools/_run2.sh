set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; tail -5 gpurun_out/r02_bench_a.err; cut -c1-3000 gpurun_out/r02_bench_a.json
TNCB_TRACE=1 timeout 300 python tools/bench_network.py --qubits 36 --rounds 10 --seed 1 --steps 2 2> gpurun_out/r02_trace_c4.txt | cut -c1-600
grep -c TRACE gpurun_out/r02_trace_c4.txt
timeout 600 python tools/sweep_engines.py 4096x4096x4096 1024x1024x1024 512x512x512 768x768x768 2048x2048x2048 4096x4096x512 4096x4096x256 65536x4096x2048 1024x1024x256 > gpurun_out/r02_sweep2.jsonl 2> gpurun_out/r02_sweep2.err; tail -3 gpurun_out/r02_sweep2.err; cat gpurun_out/r02_sweep2.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_sweep.csv python tools/sweep_engines.py 4096x4096x4096 1024x1024x1024 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_sweep.csv')) if len(r)>5]
hdr=rows[0]; print(hdr)
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:120]:
    print(r[ki][:60], r[vi])
PY
