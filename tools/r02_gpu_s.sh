#!/bin/bash
# Round 2: ncu --set full of the small kernels without a profile so far (k_halo_pack, k_init, k_design_map), raw pages as CSV.
cd "$(dirname "$0")/.."
O=gpurun_out
for k in k_halo_pack k_init k_design_map; do
  timeout 100 ncu --clock-control none --set full -k regex:$k --launch-skip 0 --launch-count 1 -f -o /tmp/r02s_$k python tools/ncu_misc.py > $O/r02s_ncu_$k.log 2>&1
  ncu -i /tmp/r02s_$k.ncu-rep --page raw --csv > $O/r02s_ncu_full_${k}_raw.csv 2>/dev/null
  ls -la $O/r02s_ncu_full_${k}_raw.csv | cut -c20-; tail -1 $O/r02s_ncu_$k.log | cut -c1-200
done
