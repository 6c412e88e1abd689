#!/bin/bash
# CTA-pair (cta_group::2) conv kernel: correctness under the forced plan, then the tiling sweep with pairs included
set +e
O=gpurun_out/s15
mkdir -p $O
AIDE_CONV_OCC=4 AIDE_CONV_TABLE=0 timeout 300 python tools/halo_probe.py --no-timing --fmts 3 2 > $O/pair_check.log 2>&1
echo "pair check rc=$?"; grep -h "==\|CHECK\|BAD\|ERR" $O/pair_check.log | cut -c1-250 | head -30; tail -3 $O/pair_check.log | cut -c1-300
if grep -q "CHECK PASS" $O/pair_check.log; then
  AIDE_CONV_OCC=4 AIDE_CONV_TABLE=0 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "conv" -p no:cacheprovider > $O/pytest_pair.log 2>&1
  echo "pytest(pair forced) rc=$?"; tail -3 $O/pytest_pair.log
  export AIDE_CONV_TABLE=0
  timeout 1100 python tools/halo_probe.py --fmts 3 --batches 8 40 --sweep-full --model fuseunet --dgrad --skip-check --skip-layers --json $O/sweep_fuse.json > $O/sweep_fuse.log 2>&1
  echo "sweep fuse rc=$?"; grep -c SWEEPF $O/sweep_fuse.log; grep -h "SWEEPF" $O/sweep_fuse.log | cut -c1-420 | head -60
fi
