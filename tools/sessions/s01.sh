#!/bin/bash
# round-2 GPU session 1: full GPU test suite, argmax audit A/B, bench, sanitizer
set +e
O=gpurun_out/s1
mkdir -p $O
nvidia-smi > $O/smi.txt 2>&1
python -m pytest tests -m gpu -q -rA --durations=25 -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/pytest.log
tail -5 $O/pytest.log
AIDE_CONV_LOSEP=0 python -m pytest tests/test_gpu_network.py -q -s -k "known_answers" -p no:cacheprovider > $O/audit_losep0.log 2>&1
AIDE_CONV_LOSEP=1 python -m pytest tests/test_gpu_network.py -q -s -k "known_answers" -p no:cacheprovider > $O/audit_losep1.log 2>&1
grep -h "256x256 parity" $O/audit_losep0.log $O/audit_losep1.log
timeout 900 python bench.py --steps 10 --warmup 3 --roofline-json $O/conv_layers.json > $O/bench.json 2> $O/bench.err
echo "bench rc=$?"; head -c 1500 $O/bench.json
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_probe.py > $O/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -3 $O/sanitizer_memcheck.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_probe.py > $O/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -3 $O/sanitizer_racecheck.log
