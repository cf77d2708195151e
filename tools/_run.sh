timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -3
for w in 0 1 2 4; do
timeout 300 python bench.py --steps 20 --no-e2e --no-cpu-baseline --no-hash-arm --dense-waves $w > gpurun_out/bench_x.json 2>gpurun_out/bench_x.err; tail -2 gpurun_out/bench_x.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_x.json')); print('c2 waves $w', d['ms_per_step'], d['roofline']['phases_ms'])"
done
for w in 0 2; do
timeout 300 python bench.py --workload c3 --steps 5 --no-cpu-baseline --no-e2e --no-hash-arm --dense-waves $w > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -2 gpurun_out/bench_c3.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_c3.json')); print('c3 waves $w', d['value'], d['ms_per_step'], d['roofline']['phases_ms'], d['parity'])"
done
