set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
timeout 300 python tools/sweep_engines.py 4096x4096x4096 65536x2048x512 32768x4096x256 1024x1024x1024 16384x4096x2048 > gpurun_out/r02_sweep5.jsonl 2> gpurun_out/r02_sweep5.err; tail -2 gpurun_out/r02_sweep5.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_sweep5.jsonl'):
    d=json.loads(l); print({k:v for k,v in d.items() if not k.endswith('_tf')})
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/r02_launches_c2b.csv python tools/sweep_engines.py 4096x4096x4096 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_c2b.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[9:16]: print(r[ki][:70], r[vi])
PY
timeout 400 python bench.py --steps 10 --warmup 3 --no-pair > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err; tail -3 gpurun_out/r02_bench_g.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_g.json') if l.startswith('{')][-1])
print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], d.get('extras_error'), 'roofline frac', d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])
print(json.dumps(d.get('sliced8_on_1gpu')))
PY
