#!/bin/bash
# One GPU-box pass: parity suite, then the single-GPU bench lines of the configs given as arguments (default: c2).
# Usage (from the repo root, under gpurun): bash tools/gpu_check.sh [tag] [workload ...]
tag=${1:-check}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
for w in "${@:-c2}"; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
  echo "bench $w exit $?"; head -c 1500 gpurun_out/${tag}_bench_$w.json; echo
done
