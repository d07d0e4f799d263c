#!/bin/bash
# round 2, GPU call AE (what is left of the budget): the reference scenes' camera parameter sets on the GPU
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 70 python -m pytest tests/test_reference_scenes_gpu.py -m gpu -q 2>&1 | tail -25 ) > gpurun_out/ae_pytest.txt
tail -25 gpurun_out/ae_pytest.txt
