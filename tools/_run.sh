timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --no-e2e --no-cpu-baseline 2> gpurun_out/bench_x.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c2', d['ms_per_step'], d['roofline']['phases_ms'], d['hash_layout']['ms_per_step'], d['hash_layout']['direct_address_with_match_cache']['ms_per_step'], d['fused_single_pass']['ms_per_step'])"
