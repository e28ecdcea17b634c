set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -3 gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n1.json') if l.startswith('{')][-1])
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], d.get('extras_error'))
print(json.dumps(d['roofline'])[:700])
print(json.dumps(d['cpu_baseline'])[:400])
print(json.dumps(d['pair_c2']['engines']), json.dumps(d['pair_c2'].get('e2e_host_buffers_pipelined')))
print(json.dumps(d.get('sliced8_on_1gpu')), json.dumps(d.get('dmma_only')), json.dumps(d['clocks']))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:crt_gemm_kernel -s 2 -c 1 -o gpurun_out/r02_crt_gemm_final -f python tools/sweep_engines.py 4096x4096x4096 > gpurun_out/ncu17.log 2>&1; tail -2 gpurun_out/ncu17.log
ncu -i gpurun_out/r02_crt_gemm_final.ncu-rep --page raw --csv > gpurun_out/r02_crt_gemm_final_raw.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-pair --no-extras --no-cpu-baseline > /dev/null 2>&1; wc -l gpurun_out/r02_launches_bench.csv
