#!/bin/bash
# tiling sweeps with the CTA-pair kernel included -> plan table -> tests + bench with the new table
set +e
O=gpurun_out/s16
mkdir -p $O
AIDE_CONV_OCC=4 AIDE_CONV_TABLE=0 timeout 600 python -m pytest tests/test_gpu_network.py -q -x -k "eval or fused" -p no:cacheprovider > $O/pytest_pair_eval.log 2>&1
echo "pytest(pair forced, eval) rc=$?"; tail -3 $O/pytest_pair_eval.log
export AIDE_CONV_TABLE=0
sweep() { name=$1; shift; timeout 1500 python tools/halo_probe.py --sweep-full --dgrad --skip-check --skip-layers --json $O/sweep_$name.json "$@" > $O/sweep_$name.log 2>&1; echo "sweep $name rc=$? $(grep -c SWEEPF $O/sweep_$name.log)"; }
sweep fuse --fmts 3 --batches 8 16 32 40 80 160 --model fuseunet
sweep unet --fmts 3 --batches 8 32 40 160 --model unet
sweep unet320 --fmts 3 --batches 8 32 40 --size 320 --model unet
sweep bf16 --fmts 2 --batches 8 40 --model fuseunet
sweep tf32 --fmts 1 --batches 8 40 --model fuseunet
unset AIDE_CONV_TABLE
python tools/make_plan_table.py $O/sweep_*_full.json
cp aide_b200/csrc/conv_plan_table.inc $O/conv_plan_table.inc
make -C aide_b200/csrc -j16 > $O/make.log 2>&1; echo "make rc=$?"
python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s16/bench.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])
for k in d:
    if k.startswith('roofline'): print(k, {a:d[k].get(a) for a in ('achieved','peak','frac','frac_sustained')})
print({k:d.get(k) for k in ('eval_single_slice','e2e_module')})
PY
AIDE_CONV_PAIR=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_nopair.json 2> $O/bench_nopair.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s16/bench_nopair.json') if l.startswith('{')][-1])
print('no pair:', {k:d.get(k) for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'])
PY
