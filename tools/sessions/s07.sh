#!/bin/bash
set +e
O=gpurun_out/s7
mkdir -p $O
nvidia-smi -L | head -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 tools/ddp_check.py > $O/ddp_check.log 2>&1
echo "ddp_check rc=$?"; grep -h "ok\|OK\|Error\|assert" $O/ddp_check.log | head -20; tail -5 $O/ddp_check.log
for BK in 4 1; do
  AIDE_B200_BUCKETS=$BK timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2963$BK bench.py --gpus 2 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/bench2_b$BK.json 2> $O/bench2_b$BK.err
  echo "bench buckets=$BK rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/s7/bench2_b$BK.json') if l.startswith('{')][-1])
    print('buckets=$BK', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'])
except Exception as e: print('parse fail', e)
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras --no-cpu-baseline --global-select > $O/bench2_global.json 2> $O/bench2_global.err
echo "bench global-select rc=$?"; tail -c 400 $O/bench2_global.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/s7/bench2_global.json') if l.startswith('{')][-1])
    print('global-select', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['engine'])
except Exception as e: print('parse fail', e)
PY
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/bench1.json 2> $O/bench1.err
python - <<PY
import json
d=json.load(open('gpurun_out/s7/bench1.json')); print('N=1', {k:d[k] for k in ('value','ms_per_step')})
PY
