#!/bin/bash
# N GPUs: weak point with and without the gradient all-reduce (AIDE_B200_NO_COMM=1: what the slowest rank's compute
# alone costs), and for N < 8 the strong point (global batch 64) and UNet-320
set +e
N=$(nvidia-smi -L | wc -l)
O=gpurun_out/s19
mkdir -p $O
run() {
  name=$1; shift
  envs=(X=1)
  while [[ "$1" == *=* ]]; do envs+=("$1"); shift; done
  if [ "$N" -gt 1 ]; then
    env "${envs[@]}" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline "$@" > $O/${name}_n$N.json 2> $O/${name}_n$N.err
  else
    env "${envs[@]}" timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu-baseline "$@" > $O/${name}_n$N.json 2> $O/${name}_n$N.err
  fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/s19/${name}_n$N.json') if l.startswith('{')][-1])
    print('$name N=$N', {k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, 'B/gpu', d['config']['per_gpu_batch'], 'e2e', d['e2e']['value'], d.get('clocks',{}).get('sm_mhz'))
except Exception as e: print('$name parse fail', e); print(open('gpurun_out/s19/${name}_n$N.err').read()[-1500:])
PY
}
run weak
if [ "$N" -eq 2 ]; then run nocomm AIDE_B200_NO_COMM=1; fi
if [ "$N" -lt 8 ]; then
  run strong --scaling strong --global-batch 64
  run unet320 --model unet --size 320
fi
