timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -3
n=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --workload c5 --steps 5 --no-cpu-baseline > gpurun_out/bench_c5_${n}gpu.json 2> gpurun_out/bench_c5_${n}gpu.err; tail -2 gpurun_out/bench_c5_${n}gpu.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/bench_c5_${n}gpu.json')); print($n, d['value'], d['ms_per_step'], d['config']['workload'], d['roofline'].get('phases_ms'), d.get('parity'))"
