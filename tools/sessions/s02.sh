#!/bin/bash
set +e
O=gpurun_out/s2
mkdir -p $O
python -m pytest tests -m gpu -q -rA -x --durations=8 -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O/pytest.log; grep -h "^C[235] \|teacher-forced\|256x256 parity" $O/pytest.log
for n in 2 1; do
  AIDE_CONV_NACC_MAX=$n python -m pytest tests/test_gpu_network.py -q -s -k "known_answers" -p no:cacheprovider > $O/audit_nacc$n.log 2>&1
  echo "NACC_MAX=$n:"; grep -h "256x256 parity\|passed\|failed" $O/audit_nacc$n.log
done
timeout 600 python tools/halo_probe.py --fmts 3 --batch 32 --sweep-full --max-cout 64 > $O/sweep_b32.log 2>&1
grep -h "SWEEPF\|ERR" $O/sweep_b32.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/bench.json 2> $O/bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/s2/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline_train_batch']['achieved'])
PY
