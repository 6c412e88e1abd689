#!/bin/bash
# wgrad on a side stream: full suite, same-session A/B of the step
set +e
O=gpurun_out/s23
mkdir -p $O
python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O/pytest.log
for v in 1 0 1 0; do
AIDE_B200_WGRAD_STREAM=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $O/bench_ws$v.json 2> $O/bench_ws$v.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/s23/bench_ws$v.json') if l.startswith('{')][-1])
print('wgrad stream $v:', {k:d.get(k) for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['clocks']['sm_mhz'])
PY
done
