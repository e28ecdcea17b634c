set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/r02_launches_c2c.csv python tools/sweep_engines.py 4096x4096x4096 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_c2c.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[9:16]: print(r[ki][:70], r[vi])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:crt_residue -s 5 -c 1 -o gpurun_out/r02_residue3 -f python tools/sweep_engines.py 4096x4096x4096 > gpurun_out/ncu15.log 2>&1; tail -2 gpurun_out/ncu15.log
ncu -i gpurun_out/r02_residue3.ncu-rep --page details 2>/dev/null | grep -E "crt_residue|Duration|Throughput|Hit Rate|Registers|Theoretical Occ|Achieved Occ|Executed Ipc|Issue Slots|Mem Busy|Max Bandwidth" | head -24
ncu -i gpurun_out/r02_residue3.ncu-rep --page raw --csv > gpurun_out/r02_residue3.raw.csv 2>/dev/null
timeout 400 python bench.py --steps 10 --warmup 3 --no-pair --no-extras --no-cpu-baseline > gpurun_out/r02_bench_h.json 2> gpurun_out/r02_bench_h.err; tail -3 gpurun_out/r02_bench_h.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_h.json') if l.startswith('{')][-1])
print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], 'roofline frac', d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])
PY
