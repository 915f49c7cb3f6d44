#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r23_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r23_pytest.log
tail -15 gpurun_out/r23_pytest.log
