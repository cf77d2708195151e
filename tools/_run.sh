timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r1b_bench_c2.json 2> gpurun_out/r1b_bench_c2.err; tail -2 gpurun_out/r1b_bench_c2.err
C2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-hash-arm"
timeout 300 $C2 > gpurun_out/plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_count_range$|^k_write_range$' -s 6 -c 2 -f -o gpurun_out/r1b_ncu_c2_range $C2 > gpurun_out/ncu.log 2>&1; tail -1 gpurun_out/ncu.log | cut -c1-80
