timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --batch 256 --steps 10 --warmup 3 --no-cpu-baseline --no-consumers 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('B256', round(d['ms_per_step'],3), d['value'], d['launches_per_step'], d['roofline'])"
