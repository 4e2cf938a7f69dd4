#!/bin/bash
timeout 300 python -m pytest tests/test_pull_gpu.py tests/test_adapter.py tests/test_replay.py tests/test_ingest_gpu.py tests/test_cabi.py -x -q -m gpu > gpurun_out/r2_v18_tests.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2_v18_tests.log
