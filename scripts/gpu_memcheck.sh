#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_plain.log 2>&1 && tail -1 gpurun_out/smoke_plain.log &&
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/memcheck.log
