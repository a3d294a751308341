run() { python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-consumers 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['launches_per_step'])"; }
run default
SCATT_ATTN_PERSIST=1 run attn_persist
SCATT_BLOCK_CLUSTER=1 run block_cluster1
SCATT_FUSED_BLOCK_MIN_TILES=1000 run no_fused_block
SCATT_L2_PREFETCH=0 run no_prefetch
run default
python tools/sweep.py batch --precision fp16x3 --batches 64,256 2>/dev/null | tail -4
SCATT_ATTN_PERSIST=1 python tools/sweep.py batch --precision fp16x3 --batches 64,256 2>/dev/null | tail -3
