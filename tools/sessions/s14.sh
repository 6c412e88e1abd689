#!/bin/bash
set +e
N=$(nvidia-smi -L | wc -l)
O=gpurun_out/s14
mkdir -p $O
run() {
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline $EXTRA > $O/${name}_n$N.json 2> $O/${name}_n$N.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/s14/${name}_n$N.json') if l.startswith('{')][-1])
    print('$name N=$N', {k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, 'e2e', d['e2e']['value'], d.get('clocks',{}).get('sm_mhz'))
except Exception as e: print('$name parse fail', e)
PY
}
EXTRA=""
run buckets4 AIDE_B200_BUCKETS=4
run buckets1 AIDE_B200_BUCKETS=1
run buckets8 AIDE_B200_BUCKETS=6
run ctas8 AIDE_B200_BUCKETS=4 NCCL_MAX_CTAS=8
run ctas16 AIDE_B200_BUCKETS=4 NCCL_MAX_CTAS=16
EXTRA="--global-select"
run globalsel AIDE_B200_BUCKETS=4
