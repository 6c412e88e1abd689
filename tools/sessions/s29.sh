#!/bin/bash
# ncu: time + DRAM bytes of every memory-side kernel of a stacked forward + train forward/backward
set +e
O=gpurun_out/s29
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"bn_|upsample2x|conv1x1|nchw|nhwc|zero_insert|adam|loss_|reduce_rows|wgrad_reduce|weight_prep|conv3x3_c3|wgrad_c3|pseudo|coteach" --csv --log-file $O/hbm_kernels.csv python tools/profile_step.py > $O/profile_step.log 2>&1
echo "ncu rc=$?"; wc -l $O/hbm_kernels.csv; tail -2 $O/profile_step.log
