#!/bin/bash
# round 2: GPU test suite (no -x: collect every failure), then smoke
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 600 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r02_pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke.log
