set -x
timeout 600 python -m pytest tests/test_gpu_tcgen05.py -x -q 2>&1 | tail -15
timeout 600 python tools/sweep_engines.py 4096x4096x4096 65536x2048x512 16384x4096x2048 4096x4096x1024 1024x1024x1024 32768x4096x256 > gpurun_out/r02_sweep_p3.jsonl 2> gpurun_out/r02_sweep_p3.err
tail -3 gpurun_out/r02_sweep_p3.err; cat gpurun_out/r02_sweep_p3.jsonl
TNCB_CRT_PRODUCTS=4 timeout 300 ncu --set full --import-source on --clock-control none -k regex:crt_gemm -s 1 -c 1 -f -o gpurun_out/r02_k512 python tools/profile_pair.py 65536x2048x512 > gpurun_out/k512.log 2>&1; tail -3 gpurun_out/k512.log
