n=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 > gpurun_out/r1b_bench_c2_${n}gpu_range.json 2> gpurun_out/r1b_bench_c2_${n}gpu_range.err; tail -2 gpurun_out/r1b_bench_c2_${n}gpu_range.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/r1b_bench_c2_${n}gpu_range.json')); print($n, round(d['value']/1e9,2), 'G/s', round(d['ms_per_step'],3), 'ms', d['config']['table_layout_chosen'], d['roofline']['phases_ms'], d['parity'], d['e2e'] and d['e2e']['ms_per_step'], d['cpu_baseline'])"
