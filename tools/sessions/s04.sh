#!/bin/bash
set +e
O=gpurun_out/s4
mkdir -p $O
python -m pytest tests -m gpu -q -x --durations=5 -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/pytest.log; grep -h "^C[235] \|teacher-forced\|256x256 parity" $O/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/bench.json 2> $O/bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/s4/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline_train_batch']['achieved'])
PY
AIDE_B200_STACK_TRAIN=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_nostack.json 2> $O/bench_nostack.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s4/bench_nostack.json'))
print('no stacked train fwd:', {k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')})
PY
export AIDE_CONV_TABLE=0
timeout 900 python tools/halo_probe.py --fmts 3 --batches 8 32 40 --sweep-full --model fuseunet --dgrad --skip-check --skip-layers --json $O/sweep_fuse.json > $O/sweep_fuse.log 2>&1
echo "sweep fuse rc=$?"; grep -c SWEEPF $O/sweep_fuse.log
timeout 900 python tools/halo_probe.py --fmts 3 --batches 8 32 --sweep-full --model unet --dgrad --skip-check --skip-layers --json $O/sweep_unet.json > $O/sweep_unet.log 2>&1
echo "sweep unet rc=$?"; grep -c SWEEPF $O/sweep_unet.log
timeout 900 python tools/halo_probe.py --fmts 3 --batches 8 32 --size 320 --sweep-full --model unet --dgrad --skip-check --skip-layers --json $O/sweep_unet320.json > $O/sweep_unet320.log 2>&1
echo "sweep unet320 rc=$?"; grep -c SWEEPF $O/sweep_unet320.log
timeout 900 python tools/halo_probe.py --fmts 2 --batches 8 32 40 --sweep-full --model fuseunet --dgrad --skip-check --skip-layers --json $O/sweep_bf16.json > $O/sweep_bf16.log 2>&1
echo "sweep bf16 rc=$?"; grep -c SWEEPF $O/sweep_bf16.log
