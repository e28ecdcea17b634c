set -x
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES_SAVE=$CUDA_VISIBLE_DEVICES
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python tools/sweep_engines.py 4096x4096x4096 65536x2048x512 32768x4096x256 1024x1024x1024 > gpurun_out/r02_sweep4.jsonl 2> gpurun_out/r02_sweep4.err; tail -2 gpurun_out/r02_sweep4.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_sweep4.jsonl'):
    d=json.loads(l); print({k:v for k,v in d.items() if not k.endswith('_tf')})
PY
for n in 4 2; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02_bench_n$n.json 2> gpurun_out/r02_bench_n$n.err; echo rc=$?; grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/r02_bench_n$n.err | tail -8
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n$n.json') if l.startswith('{')][-1])
print($n, 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'launches', d['gpu_launches'])
p=d.get('parity_n',{}); print(' same path 1gpu', p.get('same_partitioned_path_on_1gpu_ms'), 'fanin rel', p.get('fanin',{}).get('rel_diff_vs_flat'), 'sliced', p.get('sliced',{}).get('ms'), p.get('sliced',{}).get('rel_diff_vs_flat'), 'ok', p.get('ok'), d.get('extras_error'))
print(' partitioning', d['config'].get('partitioning'))
PY
done
