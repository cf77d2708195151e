python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -2
python bench.py --steps 10 --no-e2e --no-cpu-baseline > gpurun_out/bench_x.json 2>gpurun_out/bench_x.err; tail -2 gpurun_out/bench_x.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_x.json')); print(d['ms_per_step'], d['roofline']['phases_ms'], d['hash_layout']['ms_per_step'], d['hash_layout']['phases_ms'])"
python bench.py --workload c5 --steps 5 --no-cpu-baseline --no-e2e --no-hash-arm > gpurun_out/bench_r1_c5_1gpu.json 2> gpurun_out/bench_r1_c5_1gpu.err; tail -2 gpurun_out/bench_r1_c5_1gpu.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_r1_c5_1gpu.json')); print(d['value'], d['ms_per_step'], d['roofline']['phases_ms'], d['parity'])"
python bench.py --workload c4 --steps 5 --no-cpu-baseline --no-e2e --no-hash-arm > gpurun_out/bench_r1_c4.json 2> gpurun_out/bench_r1_c4.err; tail -2 gpurun_out/bench_r1_c4.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_r1_c4.json')); print(d['value'], d['ms_per_step'], d['roofline']['phases_ms'], d['parity'])"
