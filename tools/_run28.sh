set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "k2 or streaming" 2>&1 | tail -5
TNCB_TRACE=1 timeout 300 python tools/trace_slice.py bench_inputs/sycamore53_d12.json 2> gpurun_out/trace_d12_slice_v3.txt | tail -3
grep "class K2" gpurun_out/trace_d12_slice_v3.txt | sort -t' ' -k22 -n -r | awk '{print $8,$10,$12,$(NF-4),$(NF-2),$NF}' | sort -k4 -n -r | head -12
timeout 600 python tools/bench_sliced.py --path-file bench_inputs/sycamore53_d12.json --steps 2 --warmup 1 --out gpurun_out/r02_sliced_v3.jsonl 2>&1 | tail -2 | cut -c1-900
