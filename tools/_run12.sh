set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:crt_gemm_kernel -s 2 -c 1 -o gpurun_out/r02_gemm_k512 -f python tools/sweep_engines.py 65536x2048x512 > gpurun_out/ncu12a.log 2>&1; tail -2 gpurun_out/ncu12a.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"crt_residue_kernel<3>" -s 2 -c 1 -o gpurun_out/r02_residue3 -f python tools/sweep_engines.py 4096x4096x4096 > gpurun_out/ncu12b.log 2>&1; tail -2 gpurun_out/ncu12b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:crt_reconstruct -s 2 -c 1 -o gpurun_out/r02_recon -f python tools/sweep_engines.py 4096x4096x4096 > gpurun_out/ncu12c.log 2>&1; tail -2 gpurun_out/ncu12c.log
for f in r02_gemm_k512 r02_residue3 r02_recon; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; ncu -i gpurun_out/$f.ncu-rep --page details 2>/dev/null | grep -E "Duration|Throughput|Hit Rate|Registers|Theoretical Occ|Achieved Occ|Stall|Executed Ipc|Issue Slots|L1/TEX Hit|Mem Busy|Max Bandwidth|DRAM Throughput" | head -40; done
