for n in 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --no-hash-arm > gpurun_out/bench_r1_${n}gpu.json 2> gpurun_out/bench_r1_${n}gpu.err; tail -3 gpurun_out/bench_r1_${n}gpu.err | cut -c1-300; cut -c1-260 gpurun_out/bench_r1_${n}gpu.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 tools/dist_check.py 2>&1 | grep parity
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29529 bench.py --gpus 8 --workload c5 --steps 5 --no-cpu-baseline > gpurun_out/bench_r1_8gpu_c5.json 2> gpurun_out/bench_r1_8gpu_c5.err; tail -3 gpurun_out/bench_r1_8gpu_c5.err | cut -c1-300; cut -c1-400 gpurun_out/bench_r1_8gpu_c5.json
