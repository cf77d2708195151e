python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-hash-arm > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -s 64 -c 32 --csv --log-file gpurun_out/launches_c5.csv python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-hash-arm > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log | cut -c1-200
python bench.py --steps 10 --no-e2e --no-cpu-baseline > gpurun_out/bench_r1_f.json 2>gpurun_out/bench_r1_f.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_r1_f.json')); print(d['value'], d['ms_per_step'], d['roofline']['phases_ms'], d['hash_layout'])"
