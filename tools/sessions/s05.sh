#!/bin/bash
set +e
O=gpurun_out/s5
mkdir -p $O
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "spatial_attention" > $O/pytest_sa_kernel.log 2>&1
echo "sa kernel rc=$?"; tail -15 $O/pytest_sa_kernel.log
python -m pytest tests/test_gpu_network.py tests/test_gpu_trainer.py -m gpu -q -p no:cacheprovider -s -k "attention or width" > $O/pytest_sa_net.log 2>&1
echo "sa net rc=$?"; grep -h "1-cos\|passed\|failed\|Error\|error" $O/pytest_sa_net.log | head -30
python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/pytest.log
for T in 0 1; do
  AIDE_CONV_TABLE=$T timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_table$T.json 2> $O/bench_table$T.err
  python - <<PY
import json
d=json.load(open('gpurun_out/s5/bench_table$T.json'))
print('table=$T', {k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline_train_batch']['achieved'], d['clocks']['sm_mhz'])
PY
done
