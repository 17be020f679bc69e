#!/bin/bash
# ncu --set full capture of one kernel of the training step, selected by a name regex: ncu_step_kernel.sh outname regex [skip]
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" --launch-skip ${3:-3} --launch-count 1 \
  -o gpurun_out/$1 -f python scripts/launch_table.py 32 > gpurun_out/$1.log 2>&1
echo "== ncu $1 rc=$?"
