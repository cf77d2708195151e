timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --workload c3 --steps 10 --no-cpu-baseline --no-e2e --no-hash-arm > gpurun_out/r1b_bench_c3.json 2> gpurun_out/bench_c3.err; tail -2 gpurun_out/bench_c3.err; python -c "
import json,sys; d=json.load(open('gpurun_out/r1b_bench_c3.json')); print('c3', d['ms_per_step'], d['roofline']['phases_ms'], d['config']['table_layout_chosen'], d['parity'])"
timeout 600 python bench.py > gpurun_out/r1b_bench_c2.json 2> gpurun_out/r1b_bench_c2.err; tail -2 gpurun_out/r1b_bench_c2.err; python -c "
import json,sys; d=json.load(open('gpurun_out/r1b_bench_c2.json')); print('c2', d['ms_per_step'], d['roofline']['phases_ms'], d['parity'], d['e2e']['ms_per_step'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
