#!/bin/bash
# HBM-kernel round: adaptive statistics chunks, row-form upsample backward, 2x2-block upsample forward, conv1x1 backward,
# zero-padded 32-channel wgrad -- full GPU suite, probes, bench
set +e
O=gpurun_out/s20
mkdir -p $O
python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O/pytest.log
python tools/hbm_probe.py --json $O/hbm_probe.json > $O/hbm_probe.log 2>&1; grep -i "upsample\|conv1x1\|fold\|bwd" $O/hbm_probe.log
timeout 600 python tools/wgrad_probe.py --fmts 3 > $O/wgrad_probe.log 2>&1; grep -v "^\[ok" $O/wgrad_probe.log | head -30
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s20/bench.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches_per_step')}, 'e2e', d['e2e']['value'])
for k in d:
    if k.startswith('roofline'): print(k, {a:d[k].get(a) for a in ('achieved','peak','frac','frac_sustained')})
print({k:(d.get(k) or {}).get('value') for k in ('eval_single_slice','e2e_module')})
PY
