#!/bin/bash
set +e
O=gpurun_out/s11
mkdir -p $O
python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/pytest.log
python tools/hbm_probe.py --json $O/hbm_probe.json > $O/hbm_probe.log 2>&1; grep "bwd\|upsample" $O/hbm_probe.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s11/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['value'])
PY
