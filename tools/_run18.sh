set -x
mkdir -p gpurun_out
timeout 300 python tools/bench_network.py --qubits 24 --rounds 12 --seed 1 --steps 5 --resident 2>&1 | cut -c1-420
timeout 300 python tools/bench_network.py --circuit sycamore --qubits 53 --rounds 8 --seed 1 --steps 3 --path-file bench_inputs/sycamore53_d8_seed1_rg48.json --resident 2>&1 | cut -c1-420
timeout 300 python tools/bench_network.py --circuit sycamore --qubits 53 --rounds 10 --seed 1 --steps 3 --path-file bench_inputs/sycamore53_d10_seed1_rg64.json --resident 2>&1 | cut -c1-420
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tcgen05.py -x -q -k "kat or square or ragged or k3 or edge or permuted or host_pipeline" > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo memcheck rc=$?; tail -4 gpurun_out/r02_sanitizer_memcheck.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tcgen05.py -x -q -k "kat or square or k3" > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo racecheck rc=$?; grep -c "Race reported" gpurun_out/r02_sanitizer_racecheck.log; grep "Race reported\|and Read\|and Write" gpurun_out/r02_sanitizer_racecheck.log | sort | uniq -c | head; tail -3 gpurun_out/r02_sanitizer_racecheck.log
