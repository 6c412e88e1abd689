#!/bin/bash
# 2 GPUs: ddp_check after the collective-safe PeerBuffers rewrite, smoke(), one weak-scaling point
set +e
O=gpurun_out/s25
mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/ddp_check.py > $O/ddp_check.log 2>&1
echo "ddp_check rc=$?"; grep -v "^W\|^\[W\|warn\|^Setting\|^\*\*\*\|^$" $O/ddp_check.log | tail -14 | cut -c1-250
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $O/weak_n2.json 2> $O/weak_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s25/weak_n2.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['engine'].get('gradient_all_reduce'), d.get('rank_ms_per_step'))
PY
