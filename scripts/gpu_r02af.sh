#!/bin/bash
# round 2, GPU call AF (last seconds of the budget): the adaptor rebuilt against the shim that records node declarations
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 60 python -m pytest tests/test_adaptor_gpu.py -m gpu -q 2>&1 | tail -6 ) > gpurun_out/af_pytest.txt
tail -6 gpurun_out/af_pytest.txt
