n=$1
run() { # name, args...
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n "$@" --steps 10 --no-cpu-baseline > gpurun_out/r1b_${name}_${n}gpu.json 2> gpurun_out/r1b_${name}_${n}gpu.err; tail -2 gpurun_out/r1b_${name}_${n}gpu.err | cut -c1-300
  python -c "
import json; d=json.load(open('gpurun_out/r1b_${name}_${n}gpu.json')); print('$name', $n, round(d['value']/1e9,2), 'G/s', round(d['ms_per_step'],3), 'ms', d['config']['workload'][:90], d.get('parity'))"
}
if [ "$n" = "8" ]; then timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -2; fi
run bench_c2 --no-hash-arm
run bench_c5_strong --workload c5 --c5-total-log2 30
if [ "$n" != "4" ]; then run bench_c5 --workload c5; fi
