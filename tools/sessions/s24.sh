#!/bin/bash
# end-of-round evidence: full suite, argmax audit, bench, probes, conv DRAM traffic, launch list, 100-step teacher-forced run
set +e
O=gpurun_out/s24
mkdir -p $O
python -m pytest tests -m gpu -q -x -p no:cacheprovider --durations=8 > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -14 $O/pytest.log
python -m pytest tests/test_gpu_network.py -q -s -k "known_answers" -p no:cacheprovider > $O/argmax_audit.log 2>&1
grep -h "256x256\|flip\|margin" $O/argmax_audit.log | head -12
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s24/bench.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches_per_step')}, 'e2e', d['e2e']['value'], d['clocks'])
for k in d:
    if k.startswith('roofline'): print(k, {a:d[k].get(a) for a in ('achieved','peak','frac','frac_sustained','frac_of_format_ceiling')})
print({k:(d.get(k) or {}).get('value') for k in ('eval_single_slice','eval_module_single_slice','e2e_module','train_only','cpu_baseline')})
PY
python tools/hbm_probe.py --json $O/hbm_probe.json > $O/hbm_probe.log 2>&1; tail -5 $O/hbm_probe.log
timeout 600 python tools/wgrad_probe.py --fmts 3 > $O/wgrad_probe.log 2>&1; tail -2 $O/wgrad_probe.log
L="1024,512,32 128,128,128 128,128,64 128,256,32 128,64,256 256,128,128 256,256,32 256,256,64 256,512,16 32,32,256 32,64,128 512,256,64 512,512,16 512,512,32 64,128,64 64,64,128 64,64,256"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/traffic_b40.csv python tools/profile_conv.py --fmt 3 --batch 40 --reps 1 --layers $L > $O/traffic_b40.log 2>&1; echo "traffic b40 rc=$?"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/traffic_b8.csv python tools/profile_conv.py --fmt 3 --batch 8 --reps 1 --layers $L > $O/traffic_b8.log 2>&1; echo "traffic b8 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 4600 --csv --log-file $O/launches_r2k.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
echo "launch list rc=$?"; wc -l $O/launches_r2k.csv
AIDE_TF_STEPS=100 AIDE_TF_SIZE=256 timeout 1100 python -m pytest tests/test_gpu_network.py -q -s -k "teacher_forced" -p no:cacheprovider > $O/teacher_forced_100.log 2>&1
echo "teacher forced rc=$?"; tail -6 $O/teacher_forced_100.log | cut -c1-250
