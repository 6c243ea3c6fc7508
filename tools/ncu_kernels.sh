#!/bin/bash
# One `ncu --set full` capture per named kernel of the develop pipeline (run under gpurun; reports land in gpurun_out/).
# usage: tools/ncu_kernels.sh <tag> <kernel> [<kernel> ...]
tag=$1; shift
for k in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 1 -c 1 -f -o gpurun_out/${tag}_$k \
      python tools/time_develop.py --iters 1 --usm > gpurun_out/${tag}_$k.log 2>&1
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page raw --csv 2>/dev/null | python3 tools/ncu_extract.py > gpurun_out/${tag}_$k.txt
done
