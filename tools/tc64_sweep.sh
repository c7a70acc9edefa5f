#!/bin/bash
# Same-box sweep of rq_tc64_kernel's ring depths (codebook ring nb, x staging ring nx; 16 KB stages, nb + nx <= 7) through the
# Python-free harness: tells in one call whether the kernel is paced by the codebook stream, by the x stream, or by neither.
#   bash tools/tc64_sweep.sh [cluster=1|4|8]   -> gpurun_out/tc64_sweep.txt
mkdir -p gpurun_out
cl=${1:-1}
{
  timeout 40 tools/bin/tc_native_check 65536 768 3 20 /tmp/sw_ref.ids
  for nb in 2 3 4 5 6; do
    for nx in 1 2 3 4; do
      [ $((nb + nx)) -le 7 ] || continue
      echo -n "nb=$nb nx=$nx  "
      RQB200_TC_64=$cl RQB200_TC64_NB=$nb RQB200_TC64_NX=$nx timeout 40 tools/bin/tc_native_check 65536 768 3 20 /tmp/sw_var.ids
      cmp -s /tmp/sw_ref.ids /tmp/sw_var.ids || echo "    IDS DIFFER (nb=$nb nx=$nx)"
    done
  done
} > gpurun_out/tc64_sweep.txt 2>&1
cat gpurun_out/tc64_sweep.txt
