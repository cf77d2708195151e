C2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-hash-arm"
timeout 300 $C2 > gpurun_out/plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_count$|^k_write$|^k_build_dense$' -s 18 -c 6 -f -o gpurun_out/r1b_ncu_c2 $C2 > gpurun_out/ncu.log 2>&1; tail -1 gpurun_out/ncu.log | cut -c1-80
C3="python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-hash-arm"
timeout 300 $C3 > gpurun_out/plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_count_sparse$|^k_write_sparse$' -s 9 -c 3 -f -o gpurun_out/r1b_ncu_c3 $C3 > gpurun_out/ncu.log 2>&1; tail -1 gpurun_out/ncu.log | cut -c1-80
