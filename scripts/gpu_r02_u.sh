#!/bin/bash
# last sanity run of the round: full GPU suite + smoke on the final build
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r02_last.txt 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_r02_last.txt | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
