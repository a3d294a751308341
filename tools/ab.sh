timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 > gpurun_out/bench9.json 2> gpurun_out/bench9.err; tail -3 gpurun_out/bench9.err
python -c "
import json; d=json.load(open('gpurun_out/bench9.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['launches_per_step'], d['roofline']['per_kernel_ms_per_step'])"
python tools/sweep.py membound > gpurun_out/membound3.md 2>gpurun_out/membound3.err; tail -2 gpurun_out/membound3.err
