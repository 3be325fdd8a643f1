#!/bin/bash
# round 2, GPU run 37 (1 GPU): API-level tests after the last host-side changes
set -x
timeout 1200 python -m pytest tests/test_vs_reference.py tests/test_index.py tests/test_disk.py tests/test_early_stop.py tests/test_util.py -m gpu -x -q 2>&1 | tail -3
