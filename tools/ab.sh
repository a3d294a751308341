timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "linear" 2>&1 | tail -3
echo A_wide; python tools/sweep.py kernels --only linear 2>&1 | grep "fp16x3" | grep -v "^| 1600" | grep "qkv\|fc1"
python tools/sweep.py batch --precision fp16x3 --batches 64,256 2>&1 | tail -2
echo B_128; SCATT_LIB=$PWD/scattennet_b200/libscatt_b.so python tools/sweep.py kernels --only linear 2>&1 | grep "fp16x3" | grep -v "^| 1600" | grep "qkv\|fc1"
SCATT_LIB=$PWD/scattennet_b200/libscatt_b.so python tools/sweep.py batch --precision fp16x3 --batches 64,256 2>&1 | tail -2
