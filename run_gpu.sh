#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_crypto_gpu.py -q --tb=short 2>&1 | tail -40 > gpurun_out/crypto_tests.log
cat gpurun_out/crypto_tests.log
