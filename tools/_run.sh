timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --no-e2e --no-cpu-baseline --no-hash-arm > gpurun_out/bench_x.json 2>gpurun_out/bench_x.err; tail -2 gpurun_out/bench_x.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_x.json')); print('c2', d['ms_per_step'], d['roofline']['phases_ms'])"
