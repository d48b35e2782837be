#!/bin/bash
out=gpurun_out/retr; mkdir -p $out
timeout 240 python -m pytest tests/test_retrieval.py tests/test_api.py -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $out/pytest.log
timeout 120 python tools/time_retrieval.py 2>&1 | tail -3
