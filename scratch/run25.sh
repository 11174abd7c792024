#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r02_pytest25.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest25.log; grep -n "^FAILED\|^E  " gpurun_out/r02_pytest25.log | head
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke25.log 2>&1; echo "smoke rc=$?"
