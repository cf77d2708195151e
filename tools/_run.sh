set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r1b_bench_c2.json 2> gpurun_out/r1b_bench_c2.err; tail -2 gpurun_out/r1b_bench_c2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1b_bench_c2_reference_arm.json 2> gpurun_out/r1b_ref.err
for W in c3 c4 c5; do
timeout 300 python bench.py --workload $W --steps 10 --no-cpu-baseline --no-e2e --no-hash-arm > gpurun_out/r1b_bench_$W.json 2> gpurun_out/r1b_bench_$W.err; tail -2 gpurun_out/r1b_bench_$W.err
done
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r1b_launches_c2.csv $CMD > gpurun_out/ncu.log 2>&1; tail -1 gpurun_out/ncu.log | cut -c1-80
