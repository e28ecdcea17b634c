set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "forced or k3 or sliced or permute or plan_graph or host_pipeline" 2>&1 | tail -6
timeout 200 python tools/bench_permute.py > gpurun_out/r02_permute_k3.jsonl 2>&1; cat gpurun_out/r02_permute_k3.jsonl
timeout 400 python bench.py --steps 10 --warmup 3 --no-pair > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; tail -3 gpurun_out/r02_bench_e.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_e.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d.get('extras_error'))
print(json.dumps(d.get('sliced8_on_1gpu')), json.dumps(d.get('dmma_only')), json.dumps(d.get('cpu_baseline'))[:300])
PY
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tcgen05.py -x -q -k "kat or square or ragged or k3 or edge" > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo memcheck rc=$?; tail -5 gpurun_out/r02_sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tcgen05.py -x -q -k "kat or square or k3" > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo racecheck rc=$?; tail -5 gpurun_out/r02_sanitizer_racecheck.log
