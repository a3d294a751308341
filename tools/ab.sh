timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -3
run() { python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-consumers 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['launches_per_step'])"; }
for i in 1 2; do
run A_sub2
SCATT_LIB=$PWD/scattennet_b200/libscatt_b.so run B_base
done
python tools/trace_linear.py 2>&1 | sed -n 20,45p
