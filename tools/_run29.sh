set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cut -c1-600 gpurun_out/r02_bench_n1_final.json
TNCB_CRT_PRODUCTS=4 timeout 300 ncu --set full --import-source on --clock-control none -k regex:crt_gemm -s 1 -c 1 -f -o gpurun_out/r02_crt_gemm_p4 python tools/profile_pair.py 4096x4096x4096 > gpurun_out/ncu_p4.log 2>&1; tail -2 gpurun_out/ncu_p4.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:crt_gemm -s 1 -c 1 -f -o gpurun_out/r02_crt_gemm_p3 python tools/profile_pair.py 4096x4096x4096 > gpurun_out/ncu_p3.log 2>&1; tail -2 gpurun_out/ncu_p3.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pair --no-extras --no-config5 > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/bench_under_ncu.log | cut -c1-200
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_c2_final.csv python tools/profile_pair.py 4096x4096x4096 4 > /dev/null 2>&1
