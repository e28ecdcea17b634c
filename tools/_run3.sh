set -x
mkdir -p gpurun_out
timeout 120 tools/i8_peak > gpurun_out/r02_i8_peak.txt 2>&1; cat gpurun_out/r02_i8_peak.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python tools/sweep_engines.py 4096x4096x4096 1024x1024x1024 512x512x512 65536x2048x512 16384x4096x2048 > gpurun_out/r02_sweep3.jsonl 2> gpurun_out/r02_sweep3.err; tail -3 gpurun_out/r02_sweep3.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_sweep3.jsonl'):
    d=json.loads(l); print({k:v for k,v in d.items() if not k.endswith('_err') and not k.endswith('_tf')})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches_c2.csv python tools/sweep_engines.py 4096x4096x4096 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_c2.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[9:22]: print(r[ki][:70], r[vi])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:crt_gemm_kernel -s 2 -c 1 -o gpurun_out/r02_crt_gemm -f python tools/sweep_engines.py 4096x4096x4096 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ncu -i gpurun_out/r02_crt_gemm.ncu-rep --page raw --csv > gpurun_out/r02_crt_gemm_raw.csv 2>/dev/null; wc -l gpurun_out/r02_crt_gemm_raw.csv
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; tail -3 gpurun_out/r02_bench_b.err; cut -c1-1500 gpurun_out/r02_bench_b.json
