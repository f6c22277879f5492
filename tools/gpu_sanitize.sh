#!/bin/bash
# compute-sanitizer memcheck over the SIMT-heavy kernel tests (grouped conv, streamed BN, batched pack, optimizer)
mkdir -p gpurun_out
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_training.py tests/test_gpu_optim.py -m gpu -x -q -k "${1:-grouped_conv or bn_act_mask_forward or batched_weight_pack or dropout_in_bn or optim or log_softmax_argmax or entropy}" > gpurun_out/sanitize.log 2>&1; echo "sanitize rc=$?"
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/sanitize.log; grep -m5 -A12 "Invalid\|misaligned" gpurun_out/sanitize.log | cut -c1-200; tail -6 gpurun_out/sanitize.log | cut -c1-200
