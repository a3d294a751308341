timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "linear" 2>&1 | tail -3
echo A_persist; python tools/sweep.py batch --precision fp16x3 --batches 16,64,256 2>&1 | tail -3
echo B_dual; SCATT_LIB=$PWD/scattennet_b200/libscatt_b.so python tools/sweep.py batch --precision fp16x3 --batches 16,64,256 2>&1 | tail -3
