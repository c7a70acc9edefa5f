#!/bin/bash
# Python-free run of the tensor-core tokeniser (csrc/rq_tcx.cu) through the C ABI at several shapes: ids hash (fnv; the values of
# the validated kernel are in profiles/r2_tcx_bringup.txt), re-rank counters, event-timed ms, then the event timeline and the
# per-warp epilogue phase clocks of CTA 0, optionally one `ncu --set full` capture.
#   gpurun --timeout 600 -- 'bash tools/tcx_bringup.sh'      -> gpurun_out/tcx_bringup.txt, timeline_tcx.txt, ncu_tcx.ncu-rep
mkdir -p gpurun_out
{
for shape in "1 768 3" "100 768 3" "192 768 3" "193 768 3" "1000 768 3" "5000 256 4" "600 64 8" "12101 768 3" "20000 768 3" "65536 768 3" "84000 768 3"; do
  set -- $shape
  timeout 60 tools/bin/tc_native_check $1 $2 $3 20 /tmp/new.ids; echo "exit $?"
done
} > gpurun_out/tcx_bringup.txt 2>&1
RQB200_TC_TRACE=1 timeout 60 tools/bin/tc_native_check 65536 768 3 20 /tmp/tl.ids > gpurun_out/timeline_tcx.txt 2>&1
cat gpurun_out/tcx_bringup.txt; head -120 gpurun_out/timeline_tcx.txt
# one `ncu --set full` capture of the kernel at the bench shape (source-level stall reasons; read with `ncu -i ... --page source --csv`)
if [ "${TCX_NCU:-1}" = "1" ]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:rq_tcx -s 2 -c 1 -f -o gpurun_out/ncu_tcx tools/bin/tc_native_check 65536 768 3 3 /tmp/n.ids > gpurun_out/ncu_tcx.log 2>&1; echo "ncu exit $?"
fi
