timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -2
for W in c4 c5; do
timeout 300 python bench.py --workload $W --steps 5 --no-cpu-baseline --no-e2e --no-hash-arm > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; tail -2 gpurun_out/bench_$W.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$W.json')); print('$W', d['value'], d['ms_per_step'], d['roofline']['phases_ms'], d['parity'])"
done
