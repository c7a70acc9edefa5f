#!/bin/bash
# First GPU contact of rq_tc64_kernel (csrc/rq_tc64.cu): layout / rate probe, then the default kernel vs RQB200_TC_64=1 on
# the same seeded inputs through the C ABI (no Python: starts in milliseconds).  Run from the repo root on a B200 box.
mkdir -p gpurun_out
{
  timeout 20 tools/bin/pair64_probe 100 4096; echo "exit $?"
  for shape in "1000 768 3 5" "65536 768 3 20"; do
    set -- $shape
    timeout 40 tools/bin/tc_native_check $1 $2 $3 $4 /tmp/a_$1.ids; echo "exit $?"
    RQB200_TC_64=1 timeout 40 tools/bin/tc_native_check $1 $2 $3 $4 /tmp/b_$1.ids; echo "exit $?"
    cmp /tmp/a_$1.ids /tmp/b_$1.ids && echo "IDS_IDENTICAL B=$1 (clusters of 2)"
    RQB200_TC_64=4 timeout 40 tools/bin/tc_native_check $1 $2 $3 $4 /tmp/c_$1.ids; echo "exit $?"
    cmp /tmp/a_$1.ids /tmp/c_$1.ids && echo "IDS_IDENTICAL B=$1 (clusters of 4, multicast)"
    RQB200_TC_64=8 timeout 40 tools/bin/tc_native_check $1 $2 $3 $4 /tmp/d_$1.ids; echo "exit $?"
    cmp /tmp/a_$1.ids /tmp/d_$1.ids && echo "IDS_IDENTICAL B=$1 (clusters of 8, multicast)"
  done
} > gpurun_out/tc64_bringup.txt 2>&1
cat gpurun_out/tc64_bringup.txt
