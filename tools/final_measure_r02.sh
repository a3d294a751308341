# Round-2 measurement pass on one B200 (dev tool): bench lines, ncu launch list + full capture, sweeps -> gpurun_out/r02k_*
set -u
O=gpurun_out
python bench.py --steps 50 --warmup 5 > $O/r02k_bench_n1.json 2> $O/r02k_bench_n1.err; tail -2 $O/r02k_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 3 > $O/r02k_bench_ref.json 2> $O/r02k_bench_ref.err; tail -2 $O/r02k_bench_ref.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-consumers --profile-steps 1 > $O/plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/r02k_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-consumers --profile-steps 1 > $O/ncu1.log 2>&1
tail -1 $O/ncu1.log | cut -c1-120
python tools/ncu_step.py > $O/plain2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $O/r02k_step_kernels python tools/ncu_step.py > $O/ncu2.log 2>&1
tail -2 $O/ncu2.log | cut -c1-120
python tools/sweep.py batch --precision fp16x3,fp16x1 --batches 1,8,32,64,256,1024 > $O/r02k_sweep_batch.md 2> $O/sweep1.err; tail -1 $O/sweep1.err
python tools/sweep.py batch --precision fp16x3 --T 400 --config phoenix-2014 --batches 8,64,256 > $O/r02k_sweep_batch_T400.md 2> $O/sweep2.err; tail -1 $O/sweep2.err
python tools/sweep.py kernels > $O/r02k_sweep_kernels.md 2> $O/sweep3.err; tail -1 $O/sweep3.err
python tools/sweep.py membound > $O/r02k_sweep_membound.md 2> $O/sweep4.err; tail -1 $O/sweep4.err
ls -la $O/r02k_*
