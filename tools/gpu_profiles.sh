#!/bin/bash
# round-end evidence: launch list of the bench workload, full ncu of every kernel of one eager training step
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_bench_train.log 2>&1; echo "ncu-list rc=$?"
python tools/summarize_launches.py gpurun_out/launches_train.csv > gpurun_out/launches_train_summary.md; head -30 gpurun_out/launches_train_summary.md
timeout 1500 ncu --set full --clock-control none -k regex:'absmax|logmel|instnorm|conv1d_umma|wgrad_umma|bn_stream|bn_finalize|pack_weight|unpack_wgrad|mt_|bct_to_btc|log_softmax|ctc_' -s 534 -c 178 -f -o /tmp/prof_all python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-cuda-graphs > gpurun_out/ncu_all.log 2>&1; echo "ncu-full rc=$?"
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_train_all_raw.csv 2>/dev/null; ls -la gpurun_out/prof_train_all_raw.csv
