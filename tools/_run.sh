timeout 300 python bench.py --workload c3 --steps 3 --no-cpu-baseline --no-e2e --no-hash-arm 2> gpurun_out/bench_c3.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c3', d['ms_per_step'], d['config']['table_layout_chosen'], d['roofline']['kernel'][:40])"
timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-e2e --no-hash-arm 2> gpurun_out/bench_x.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c2', d['ms_per_step'], d['config']['table_layout_chosen'], d['roofline']['frac'], d['roofline']['job']['frac'])"
timeout 300 python bench.py --workload c4 --steps 3 --no-cpu-baseline --no-e2e --no-hash-arm 2> gpurun_out/bench_c4.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c4', d['ms_per_step'], d['config']['table_layout_chosen'])"
tail -2 gpurun_out/bench_c3.err gpurun_out/bench_x.err gpurun_out/bench_c4.err | grep -i "error\|Traceback" | head
