#!/bin/bash
# pair kernel: new test, compute-sanitizer, ncu --set full on the layers VERDICT names, launch list of a step
set +e
O=gpurun_out/s18
mkdir -p $O
python -m pytest tests/test_gpu_kernels.py -q -x -k "cta_pair" -p no:cacheprovider > $O/pytest_pair.log 2>&1
echo "pytest pair rc=$?"; tail -5 $O/pytest_pair.log
AIDE_CONV_OCC=4 AIDE_CONV_TABLE=0 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_probe.py > $O/sanitizer_memcheck_pair.log 2>&1
echo "memcheck rc=$?"; tail -3 $O/sanitizer_memcheck_pair.log
AIDE_CONV_OCC=4 AIDE_CONV_TABLE=0 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_probe.py > $O/sanitizer_racecheck_pair.log 2>&1
echo "racecheck rc=$?"; tail -3 $O/sanitizer_racecheck_pair.log
AIDE_CONV_OCC=4 AIDE_CONV_TABLE=0 timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 python tools/sanitize_probe.py > $O/sanitizer_synccheck_pair.log 2>&1
echo "synccheck rc=$?"; tail -3 $O/sanitizer_synccheck_pair.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -c 16 -o $O/conv_pair_full_b40 python tools/profile_conv.py --fmt 3 --batch 40 --reps 2 --layers 128,64,256 64,64,256 32,32,256 64,64,128 32,64,128 512,256,64 256,128,128 1024,512,32 > $O/profile_conv_b40.log 2>&1
echo "ncu conv b40 rc=$?"; tail -12 $O/profile_conv_b40.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -c 10 -o $O/conv_pair_full_b8 python tools/profile_conv.py --fmt 3 --batch 8 --reps 2 --layers 128,64,256 64,64,256 32,32,256 64,64,128 512,256,64 > $O/profile_conv_b8.log 2>&1
echo "ncu conv b8 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 4600 --csv --log-file $O/launches_r2i.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
echo "launch list rc=$?"; wc -l $O/launches_r2i.csv
ls -la $O
