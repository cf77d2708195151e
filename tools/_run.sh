python -m pytest tests -m gpu -x -q -k "golden or c1 or ragged or heavy or i64" 2>&1 | tail -2
for w in c2 c4; do python bench.py --workload $w --steps 8 --no-e2e --no-cpu-baseline --no-hash-arm > gpurun_out/bench_r1_$w.json 2>gpurun_out/bench_r1_$w.err; tail -2 gpurun_out/bench_r1_$w.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_r1_$w.json')); print('$w', d['value'], d['ms_per_step'], d['roofline']['phases_ms'], d['roofline']['job']['frac'], d['parity'])"; done
