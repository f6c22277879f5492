#!/bin/bash
mkdir -p gpurun_out
echo "== old"; CONVASR_B200_LIB=/root/repo/lib_old.so timeout 300 python tools/time_train_kernels.py 2>&1 | tail -6
echo "== new"; timeout 300 python tools/time_train_kernels.py 2>&1 | tail -6
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -x -q 2>&1 | tail -3
