python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; tail -2 gpurun_out/bench_r1_final.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_final.json')); print(d['value'], d['ms_per_step'], d['roofline']['phases_ms'], d['roofline']['frac'], d['roofline']['job']['frac'], d['e2e']['ms_per_step'], d['clocks'], d['hash_layout']['ms_per_step'])"
