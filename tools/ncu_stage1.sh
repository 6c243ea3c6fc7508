#!/bin/bash
for k in k_hsl k_box_h_tiles k_box_v k_teq_apply; do
  ncu --set full --clock-control none -k regex:"^$k" -s 2 -c 1 -f -o gpurun_out/r2s4_$k python tools/time_stage1.py > gpurun_out/r2s4_$k.log 2>&1
  ncu -i gpurun_out/r2s4_$k.ncu-rep --page raw --csv 2>/dev/null | python3 tools/ncu_extract.py > gpurun_out/r2s4_ncu_$k.txt
  rm -f gpurun_out/r2s4_$k.ncu-rep
done
tail -n 20 gpurun_out/r2s4_ncu_k_hsl.txt
