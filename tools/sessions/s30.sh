#!/bin/bash
# ncu --set full on the first-layer kernel (what bounds it?)
set +e
O=gpurun_out/s30
mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_c3_fwd -c 2 -o $O/c3_full python tools/profile_step.py > $O/c3.log 2>&1
echo "ncu rc=$?"; ls -la $O
