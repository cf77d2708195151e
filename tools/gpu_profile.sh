#!/bin/bash
# ncu --set full captures of the dominant kernels (one GPU): config 2 default, config 2 with sparse keys, config 5-shape.
# Usage (under gpurun): bash tools/gpu_profile.sh <tag>
tag=${1:-prof}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline --no-e2e"
$B --workload c2 > gpurun_out/${tag}_plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_write_range|k_count_range|k_build_dense" -s 9 -c 3 -o gpurun_out/${tag}_c2 $B --workload c2 > gpurun_out/${tag}_ncu_c2.log 2>&1; echo "c2 $?"
$B --workload c2s > gpurun_out/${tag}_plain_c2s.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"^k_count$|^k_write$|k_build_hash|k_rp_scatter|k_rp_hist" -s 18 -c 6 -o gpurun_out/${tag}_c2s $B --workload c2s > gpurun_out/${tag}_ncu_c2s.log 2>&1; echo "c2s $?"
$B --workload c5 > gpurun_out/${tag}_plain_c5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_rp_scatter|k_rp_hist|k_rj_join|k_rj_emit" -s 33 -c 11 -o gpurun_out/${tag}_c5 $B --workload c5 > gpurun_out/${tag}_ncu_c5.log 2>&1; echo "c5 $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${tag}_launches_c2s.csv $B --workload c2s > /dev/null 2>&1; echo "launches c2s $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${tag}_launches_c4.csv $B --workload c4 > /dev/null 2>&1; echo "launches c4 $?"
