#!/bin/bash
set +e
O=gpurun_out/s8
mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 4200 --csv --log-file $O/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
echo "launch list rc=$?"; wc -l $O/launches_r2.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -c 12 -o $O/conv_full_r2 python tools/profile_conv.py --fmt 3 --batch 32 --reps 2 --layers 128,64,256 64,64,256 32,32,256 64,64,128 512,256,64 256,128,128 > $O/profile_conv.log 2>&1
echo "ncu conv rc=$?"; ls -la $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_ -s 30 -c 8 -o $O/wgrad_full_r2 python tools/wgrad_probe.py --fmts 3 > $O/profile_wgrad.log 2>&1
echo "ncu wgrad rc=$?"; tail -3 $O/profile_wgrad.log
