set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -22
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err; tail -3 gpurun_out/r02_bench_f.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_f.json') if l.startswith('{')][-1])
print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['wall_ms_per_step'], 'launches', d['gpu_launches'], d.get('extras_error'))
print(json.dumps(d.get('sliced8_on_1gpu')), json.dumps(d.get('roofline'))[:400])
print(json.dumps(d['pair_c2']['engines']), json.dumps(d['pair_c2'].get('e2e_host_buffers_pipelined')))
PY
timeout 200 python tools/bench_network.py --qubits 20 --rounds 8 --seed 4 --steps 30 --resident 2>&1 | cut -c1-400
timeout 200 python tools/bench_network.py --qubits 20 --rounds 8 --seed 4 --steps 30 2>&1 | cut -c1-400
