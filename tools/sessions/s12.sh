#!/bin/bash
# 2 GPUs: NCCL CTA budget vs bucketed all-reduce; strong scaling point N=2 (global 64 -> 32 per GPU)
set +e
O=gpurun_out/s12
mkdir -p $O
run() { # name, env..., args
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus 2 --steps 10 --warmup 3 --no-extras --no-cpu-baseline $EXTRA > $O/$name.json 2> $O/$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/s12/$name.json') if l.startswith('{')][-1])
    print('$name', {k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['config']['per_gpu_batch'], d['e2e']['value'])
except Exception as e: print('$name parse fail', e)
PY
}
EXTRA=""
run base_b4 AIDE_B200_BUCKETS=4
run ctas2_b4 AIDE_B200_BUCKETS=4 NCCL_MAX_CTAS=2
run ctas4_b4 AIDE_B200_BUCKETS=4 NCCL_MAX_CTAS=4
run ctas8_b4 AIDE_B200_BUCKETS=4 NCCL_MAX_CTAS=8
run ctas4_b1 AIDE_B200_BUCKETS=1 NCCL_MAX_CTAS=4
run ctas8_b1 AIDE_B200_BUCKETS=1 NCCL_MAX_CTAS=8
EXTRA="--scaling strong --global-batch 64"
run strong_n2 AIDE_B200_BUCKETS=4
EXTRA=""
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/n1.json 2> $O/n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s12/n1.json')); print('N=1', {k:d[k] for k in ('value','ms_per_step')})
PY
