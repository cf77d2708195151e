#!/bin/bash
# One GPU-box pass: parity suite, the default bench line (headline + extras), then per-launch times (ncu) of the workloads given.
# Usage (from the repo root, under gpurun): bash tools/gpu_check.sh <tag> [workload ...]
tag=${1:-check}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; tail -3 gpurun_out/${tag}_bench.err
for w in "$@"; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${tag}_launches_$w.csv \
    python bench.py --workload $w --steps 1 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_$w.log 2>&1
  echo "ncu $w exit $?"
done
