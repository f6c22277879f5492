#!/bin/bash
# sweep of the streamed BatchNorm kernels' CTA size / stage budget (tools/time_train_kernels.py, one process per setting)
mkdir -p gpurun_out
: > gpurun_out/r02b_bn_geometry.log
for th in 128 256 512; do for kb in 32 48 72 96; do
	CONVASR_B200_BN_THREADS=$th CONVASR_B200_BN_STAGE_KB=$kb timeout 60 python tools/time_train_kernels.py 2>&1 | grep "^\[" | tee -a gpurun_out/r02b_bn_geometry.log | cut -c1-140
done; done
