# A/B of two builds of libscatt.so inside ONE gpurun call (dev tool).  Box-to-box variation between calls is ~1 %,
# more than most single optimisations move the B=8 step, so variants are only ever compared back to back on one box:
#
#   python - <<'PY'                              # build the variant next to the default library (here, on CPU)
#   from scattennet_b200 import build as B
#   B.build(extra_flags=("-DSCATT_SUB2=0",), out="scattennet_b200/libscatt_b.so")   # compile-time switches: SCATT_SUB2,
#   PY                                                                              # SCATT_PERSIST, SCATT_PERSIST_WIDE, SCATT_RES_STAGED
#   gpurun --timeout 900 -- 'bash tools/ab.sh'
#
# SCATT_LIB selects the library a process loads (scattennet_b200/_lib.py).
run() { python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-consumers 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['launches_per_step'])"; }
for i in 1 2; do
  run A_default
  [ -f scattennet_b200/libscatt_b.so ] && SCATT_LIB=$PWD/scattennet_b200/libscatt_b.so run B_variant
done
