#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fastslam" 2>&1 | tail -4
