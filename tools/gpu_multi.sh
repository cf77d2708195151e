#!/bin/bash
# Multi-GPU pass (under gpurun --gpus N): both plans against the oracle (tools/dist_check.py through tests/test_gpu_dist.py), then the
# bench line at N GPUs: headline = config 2 broadcast build (weak), extras c3 (strong) and c5 (radix plan, exchange fused, weak).
# Usage: bash tools/gpu_multi.sh <tag> <N> [extra bench args]
tag=$1; n=$2; shift 2
mkdir -p gpurun_out
python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/${tag}_dist_pytest.log 2>&1; echo "dist pytest exit $?" >> gpurun_out/${tag}_dist_pytest.log
tail -4 gpurun_out/${tag}_dist_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 "$@" \
  > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
echo "bench exit $?"; tail -3 gpurun_out/${tag}_bench_${n}gpu.err; head -c 600 gpurun_out/${tag}_bench_${n}gpu.json
# what the host side gives all ranks at once (ceiling of the e2e leg), with and without the NUMA binding
for b in "" "--no-bind"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 tools/pcie_probe.py $b \
    >> gpurun_out/${tag}_pcie_${n}gpu.jsonl 2>> gpurun_out/${tag}_pcie_${n}gpu.err
done
tail -2 gpurun_out/${tag}_pcie_${n}gpu.jsonl | cut -c1-400
