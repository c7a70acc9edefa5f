#!/bin/bash
# First GPU contact of the kernel variants written without GPU access at the end of round 1: the
# layout / rate probe, then the default kernel vs every variant on the same seeded inputs through the C ABI (no Python:
# starts in milliseconds).  Run from the repo root on a B200 box.  ~20 s of GPU time.
mkdir -p gpurun_out
{
  timeout 20 tools/bin/pair64_probe 100 4096; echo "exit $?"
  for shape in "1000 768 3 5" "65536 768 3 20"; do
    set -- $shape
    timeout 40 tools/bin/tc_native_check $1 $2 $3 $4 /tmp/ref_$1.ids; echo "exit $?"
    for v in "RQB200_TC_FASTSCAN=1" "RQB200_TC_TMA=1" "RQB200_TC_TMA=1 RQB200_TC_FASTSCAN=1" "RQB200_TC_TMA=1 RQB200_TC_PAIR=1" "RQB200_TC_TMA=1 RQB200_TC_PAIR=1 RQB200_TC_FASTSCAN=1" "RQB200_TC_PAIR=1 RQB200_TC_FASTSCAN=1" "RQB200_TC_64=1" "RQB200_TC_64=1 RQB200_TC64_GROUPS=2" "RQB200_TC_64=4" "RQB200_TC_64=4 RQB200_TC64_GROUPS=2" "RQB200_TC_64=8"; do
      env $v timeout 40 tools/bin/tc_native_check $1 $2 $3 $4 /tmp/var_$1.ids; echo "exit $?"
      cmp /tmp/ref_$1.ids /tmp/var_$1.ids && echo "IDS_IDENTICAL B=$1 [$v]"
      rm -f /tmp/var_$1.ids
    done
  done
} > gpurun_out/tc64_bringup.txt 2>&1
cat gpurun_out/tc64_bringup.txt
