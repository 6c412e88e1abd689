#!/bin/bash
set +e
O=gpurun_out/s3
mkdir -p $O
# correctness of the resident-weight plans + new nacc rule first
AIDE_CONV_WRES=1 python tools/sanitize_probe.py > $O/probe_wres1.log 2>&1; echo "probe wres=1 rc=$?"; tail -2 $O/probe_wres1.log
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py -m gpu -q -x -p no:cacheprovider -k "conv or known_answers or golden" > $O/pytest_conv.log 2>&1
echo "pytest rc=$?"; tail -3 $O/pytest_conv.log; grep -h "256x256 parity" $O/pytest_conv.log
for B in 32 8; do
  timeout 900 python tools/halo_probe.py --fmts 3 --batch $B --sweep-full --model both --skip-check --skip-layers --json $O/sweep_b$B.json > $O/sweep_b$B.log 2>&1
  echo "sweep B=$B rc=$?"; grep -c SWEEPF $O/sweep_b$B.log; grep ERR $O/sweep_b$B.log | head -5
done
timeout 600 python tools/halo_probe.py --fmts 3 --batch 8 --size 320 --sweep-full --model unet --skip-check --skip-layers --json $O/sweep_u320_b8.json > $O/sweep_u320_b8.log 2>&1
timeout 600 python tools/halo_probe.py --fmts 3 --batch 32 --size 320 --sweep-full --model unet --skip-check --skip-layers --json $O/sweep_u320_b32.json > $O/sweep_u320_b32.log 2>&1
timeout 600 python tools/halo_probe.py --fmts 2 --batch 32 --sweep-full --model fuseunet --skip-check --skip-layers --json $O/sweep_bf16_b32.json > $O/sweep_bf16_b32.log 2>&1
timeout 600 python tools/halo_probe.py --fmts 2 --batch 8 --sweep-full --model fuseunet --skip-check --skip-layers --json $O/sweep_bf16_b8.json > $O/sweep_bf16_b8.log 2>&1
ls -la $O
