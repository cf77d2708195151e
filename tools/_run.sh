timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --workload c4 --steps 5 --no-cpu-baseline --no-e2e --no-hash-arm > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -2 gpurun_out/bench_c4.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_c4.json')); print('c4', d['value'], d['ms_per_step'], d['roofline']['phases_ms'], d['parity'])"
( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real; tail -2 gpurun_out/bench_default.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real; tail -2 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | cut -c1-600
