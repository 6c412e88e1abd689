#!/bin/bash
set +e
O=gpurun_out/s10
mkdir -p $O
python tools/wgrad_probe.py --fmts 3 > $O/wgrad_probe.log 2>&1; echo "wgrad probe rc=$?"; grep -v "^\[ok" $O/wgrad_probe.log | tail -30
python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/pytest.log
python tools/hbm_probe.py > $O/hbm_probe.log 2>&1; grep "bwd_apply" $O/hbm_probe.log
timeout 900 python bench.py --steps 10 --warmup 3 --roofline-json $O/conv_layers.json > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s10/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac_of_format_ceiling'], d['roofline_train_batch']['achieved'], d['roofline_wgrad']['achieved'], d['train_only']['value'], d['fast_mode'])
PY
AIDE_WGRAD_STACKM=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_nostackm.json 2> $O/bench_nostackm.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s10/bench_nostackm.json'))
print('no STACKM wgrad', {k:d[k] for k in ('value','ms_per_step')})
PY
