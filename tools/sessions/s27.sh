#!/bin/bash
# final confirmation after the last host-side changes: full GPU suite, smoke, short bench
set +e
O=gpurun_out/s27
mkdir -p $O
python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s27/bench.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches_per_step')}, 'e2e', d['e2e']['value'], d['roofline']['achieved'], d['roofline'].get('traffic'))
PY
