python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in c2 c3 c4; do python bench.py --workload $w --steps 8 --no-e2e --no-cpu-baseline > gpurun_out/bench_r1_$w.json 2>gpurun_out/bench_r1_$w.err; tail -2 gpurun_out/bench_r1_$w.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_r1_$w.json')); print('$w', d['value'], d['ms_per_step'], d['roofline']['phases_ms'], d['roofline']['job']['frac'], d['parity'], d['hash_layout']['ms_per_step'], d['hash_layout']['phases_ms'])"; done
