#!/bin/bash
# checkpoint / resume of the fused trainer
set +e
O=gpurun_out/s28
mkdir -p $O
python -m pytest tests/test_gpu_trainer.py -q -x -p no:cacheprovider -k "checkpoint" > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -12 $O/pytest.log | cut -c1-300
