#!/bin/bash
# first-layer kernel (2 pixels per thread), batch-1 tiling sweeps for the single-slice evaluation path, full suite + bench
set +e
O=gpurun_out/s22
mkdir -p $O
python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O/pytest.log
python tools/hbm_probe.py --json $O/hbm_probe.json > $O/hbm_probe.log 2>&1; grep -i "first layer\|fold\|upsample\|conv1x1_bwd\|bwd_apply" $O/hbm_probe.log
export AIDE_CONV_TABLE=0
timeout 600 python tools/halo_probe.py --sweep-full --skip-check --skip-layers --json $O/sweep_fuse_b1.json --fmts 3 --batches 1 --model fuseunet > $O/sweep_fuse_b1.log 2>&1; echo "sweep b1 fuse rc=$? $(grep -c SWEEPF $O/sweep_fuse_b1.log)"
timeout 600 python tools/halo_probe.py --sweep-full --skip-check --skip-layers --json $O/sweep_unet_b1.json --fmts 3 --batches 1 --model unet > $O/sweep_unet_b1.log 2>&1; echo "sweep b1 unet rc=$?"
timeout 600 python tools/halo_probe.py --sweep-full --skip-check --skip-layers --json $O/sweep_unet320_b1.json --fmts 3 --batches 1 --size 320 --model unet > $O/sweep_unet320_b1.log 2>&1; echo "sweep b1 unet320 rc=$?"
unset AIDE_CONV_TABLE
grep -h SWEEPF $O/sweep_fuse_b1.log | cut -c1-260 | head -30
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s22/bench.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches_per_step')}, 'e2e', d['e2e']['value'])
for k in d:
    if k.startswith('roofline'): print(k, {a:d[k].get(a) for a in ('achieved','peak','frac','frac_sustained')})
print({k:(d.get(k) or {}).get('value') for k in ('eval_single_slice','e2e_module')})
PY
