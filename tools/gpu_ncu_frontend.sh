#!/bin/bash
# launch list + full capture of the frontend kernels (one forward of tools/time_frontend.py's shape)
mkdir -p gpurun_out
cat > /tmp/fe_once.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from convasr_b200 import models
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
sig = (torch.randn(80, 120000, generator = g) * 3000).round().clamp(-32767, 32767).to(torch.int16).to(dev)
xlen = torch.ones(80, device = dev)
fe = models.LogFilterBankFrontend(64, 8000, .02, .01, 'hann_window').to(dev)
norm = models.MaskedInstanceNorm1d(64, affine = False, eps = 2.0**-14, track_running_stats = False, temporal_mask = True, legacy = True)
for _ in range(3):
	fe.features(sig, xlen, norm, True, 64, False)
torch.cuda.synchronize()
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_frontend_launches.csv python /tmp/fe_once.py > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r02_frontend_launches.csv')) if len(r) > 10]
hdr = rows[0]; ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
for r in rows[1:]:
    print(r[ki][:60], r[vi])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:logmel16 -s 2 -c 1 -o gpurun_out/r02_logmel16 python /tmp/fe_once.py > /dev/null 2>&1; ls -la gpurun_out/r02_logmel16.ncu-rep
