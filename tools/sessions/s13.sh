#!/bin/bash
# N GPUs (N = number visible): weak scaling point, UNet-320 kidney flavour (BASELINE config 5), strong scaling point
set +e
N=$(nvidia-smi -L | wc -l)
O=gpurun_out/s13
mkdir -p $O
run() {
  name=$1; shift
  if [ "$N" -gt 1 ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-cpu-baseline "$@" > $O/${name}_n$N.json 2> $O/${name}_n$N.err
  else
    timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-extras --no-cpu-baseline "$@" > $O/${name}_n$N.json 2> $O/${name}_n$N.err
  fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/s13/${name}_n$N.json') if l.startswith('{')][-1])
    print('$name N=$N', {k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, 'B/gpu', d['config']['per_gpu_batch'], 'e2e', d['e2e']['value'], d.get('clocks',{}).get('sm_mhz'))
except Exception as e: print('$name parse fail', e); import subprocess; print(open('gpurun_out/s13/${name}_n$N.err').read()[-1500:])
PY
}
run weak
run unet320 --model unet --size 320
if [ "$N" -lt 8 ]; then run strong --scaling strong --global-batch 64; fi
