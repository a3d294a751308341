SCATT_LIB=$PWD/scattennet_b200/libscatt_c.so timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attention" 2>&1 | tail -2
run() { python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-consumers 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['launches_per_step'])"; }
for i in 1 2; do
run A_packed
SCATT_LIB=$PWD/scattennet_b200/libscatt_c.so run C_poly
done
SCATT_LIB=$PWD/scattennet_b200/libscatt_c.so python tools/trace_attention.py 2>&1 | grep -A11 "fp16x3 kind=[01]" 
