for n in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --workload c5 --c5-total-log2 30 --steps 5 --no-cpu-baseline > gpurun_out/bench_r1_c5_strong_${n}gpu.json 2> gpurun_out/bench_r1_c5_strong_${n}gpu.err; tail -2 gpurun_out/bench_r1_c5_strong_${n}gpu.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_c5_strong_${n}gpu.json')); print($n, d['value'], d['ms_per_step'], d['config']['workload'])"
done
