#!/bin/bash
cd $GRAFT_REPO_ROOT
( time python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" ) > gpurun_out/r27_smoke.log 2>&1
tail -5 gpurun_out/r27_smoke.log
( time python bench.py --impl reference ) > gpurun_out/r27_ref.json 2> gpurun_out/r27_ref.err
tail -3 gpurun_out/r27_ref.err; cat gpurun_out/r27_ref.json | cut -c1-700
( time python bench.py ) > gpurun_out/r27_default.json 2> gpurun_out/r27_default.err
tail -4 gpurun_out/r27_default.err; cat gpurun_out/r27_default.json | cut -c1-400
