#!/bin/bash
# learned_bilinear=True inside the fused trainer: new tests + the suites that touch the gradient layout
set +e
O=gpurun_out/s26
mkdir -p $O
python -m pytest tests/test_gpu_trainer.py tests/test_gpu_network.py -q -x -p no:cacheprovider -k "learned_bilinear or trainer_step_vs_oracle or graph_replay or kidney or attention" > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -12 $O/pytest.log | cut -c1-300
