# A/B of one run-time switch inside ONE gpurun call (dev tool): bash tools/ab_env.sh SCATT_FRONTEND_TC 1 0
VAR=$1; A=$2; B=$3
run() { env $VAR=$2 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-consumers 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['launches_per_step'])"; }
for i in 1 2; do
  run "$VAR=$A" $A
  run "$VAR=$B" $B
done
