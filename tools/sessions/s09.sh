#!/bin/bash
set +e
O=gpurun_out/s9
mkdir -p $O
python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/pytest.log
python tools/hbm_probe.py --json $O/hbm_probe.json > $O/hbm_probe.log 2>&1; cat $O/hbm_probe.log | tail -40
export AIDE_CONV_TABLE=0
timeout 900 python tools/halo_probe.py --fmts 3 --batches 8 40 --sweep-full --model fuseunet --dgrad --skip-check --skip-layers --json $O/sweep_fuse.json > $O/sweep_fuse.log 2>&1
echo "sweep fuse rc=$?"; grep -c SWEEPF $O/sweep_fuse.log; grep -h "SWEEPF" $O/sweep_fuse.log | grep " 40\b\|@256\|@128" | cut -c1-330 | head -24
timeout 900 python tools/halo_probe.py --fmts 3 --batches 8 32 --sweep-full --model unet --dgrad --skip-check --skip-layers --json $O/sweep_unet.json > $O/sweep_unet.log 2>&1
timeout 900 python tools/halo_probe.py --fmts 3 --batches 8 32 --size 320 --sweep-full --model unet --dgrad --skip-check --skip-layers --json $O/sweep_unet320.json > $O/sweep_unet320.log 2>&1
timeout 900 python tools/halo_probe.py --fmts 2 --batches 8 40 --sweep-full --model fuseunet --dgrad --skip-check --skip-layers --json $O/sweep_bf16.json > $O/sweep_bf16.log 2>&1
unset AIDE_CONV_TABLE
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s9/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline_train_batch']['achieved'])
PY
