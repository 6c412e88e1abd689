#!/bin/bash
# N GPUs: peer-memory all-reduce -- semantics check (tools/ddp_check.py at N = 2), then weak-scaling bench p2p vs nccl
set +e
N=$(nvidia-smi -L | wc -l)
O=gpurun_out/s21
mkdir -p $O
if [ "$N" -ne 2 ]; then export DDP_CHECK_KERNEL_ONLY=1; fi
if [ "$N" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/ddp_check.py > $O/ddp_check.log 2>&1
  echo "ddp_check rc=$?"; grep -v "^W\|^\[W\|warn" $O/ddp_check.log | tail -22 | cut -c1-300
fi
unset DDP_CHECK_KERNEL_ONLY
run() {
  name=$1; shift
  envs=(X=1)
  while [[ "$1" == *=* ]]; do envs+=("$1"); shift; done
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline "$@" > $O/${name}_n$N.json 2> $O/${name}_n$N.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/s21/${name}_n$N.json') if l.startswith('{')][-1])
    print('$name N=$N', {k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, 'e2e', d['e2e']['value'], d.get('clocks',{}).get('sm_mhz'), d.get('rank_ms_per_step'))
except Exception as e: print('$name parse fail', e); print(open('gpurun_out/s21/${name}_n$N.err').read()[-2500:])
PY
}
run p2p AIDE_B200_COMM=p2p
run synconly AIDE_B200_COMM=p2p AIDE_B200_COMM_SYNC_ONLY=1
run nocomm AIDE_B200_NO_COMM=1
run nccl AIDE_B200_COMM=nccl
