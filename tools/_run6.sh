set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "forced or k3 or host_pipeline or plan_graph or sliced or permute or engine_is_really or statevector" 2>&1 | tail -8
timeout 200 python tools/bench_permute.py > gpurun_out/r02_permute_k3.jsonl 2>&1; cat gpurun_out/r02_permute_k3.jsonl
TNCB_NO_K3=1 timeout 200 python tools/bench_permute.py > gpurun_out/r02_permute_plain.jsonl 2>&1; cat gpurun_out/r02_permute_plain.jsonl
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; tail -3 gpurun_out/r02_bench_d.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_d.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])
print(json.dumps(d['pair_c2']['engines']['tcgen05_modular']), json.dumps(d['pair_c2'].get('e2e_host_buffers')), json.dumps(d['pair_c2'].get('e2e_host_buffers_pipelined')))
PY
