#!/bin/bash
# step_from_host with a prefetched next batch: equality test, then the bench's e2e leg
set +e
O=gpurun_out/s31
mkdir -p $O
python -m pytest tests/test_gpu_trainer.py -q -x -p no:cacheprovider -k "prefetch or graph_replay" > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -8 $O/pytest.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s31/bench.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
PY
