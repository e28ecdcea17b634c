set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 240 python -m pytest tests/test_gpu_tcgen05.py -x -q -k "engine_is_really or square" 2>&1 | tail -15
echo "== quick done rc=$?"
timeout 900 python -m pytest tests/test_gpu_tcgen05.py -x -q 2>&1 | tail -25
timeout 600 python tools/sweep_engines.py 4096x4096x4096 1024x1024x1024 512x512x512 4096x4096x512 65536x2048x512 256x128x262144 > gpurun_out/r02_sweep1.jsonl 2> gpurun_out/r02_sweep1.err; tail -3 gpurun_out/r02_sweep1.err; cat gpurun_out/r02_sweep1.jsonl
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
