#!/bin/bash
set +e
O=gpurun_out/s6
mkdir -p $O
python -m pytest tests -m gpu -q -x -p no:cacheprovider -s > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/pytest.log; grep -h "eval forward launches\|single-slice\|Error" $O/pytest.log | head
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/s6/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline_train_batch']['achieved'])
for k in ('eval_single_slice','e2e_module','train_only','roofline_wgrad','roofline_hbm','cpu_baseline'):
    v=d.get(k); 
    if isinstance(v,dict): v={a:b for a,b in v.items() if a not in ('layers','api','note','sample','kernel')}
    print(k, v)
PY
