#!/bin/bash
# The first GPU call of round 2 (about a minute of B200 time, no Python): everything DESIGN.md 5.2d says is needed to decide
# what bounds the tokeniser -- L2->SM streaming ceiling, event timelines of every variant, ring-depth sweep.
#   gpurun --timeout 300 -- 'bash tools/r2_first_call.sh'       (tools/bin/* and librqb200.so are built in the container first)
mkdir -p gpurun_out
timeout 60 tools/bin/l2_stream_probe > gpurun_out/l2_stream_probe.txt 2>&1; echo "l2 probe exit $?"
bash tools/tc64_bringup.sh > /dev/null 2>&1; echo "bring-up exit $?"; grep -c IDS_IDENTICAL gpurun_out/tc64_bringup.txt
bash tools/tc_timelines.sh
bash tools/tc64_sweep.sh 1 > /dev/null 2>&1; echo "sweep exit $?"
tail -20 gpurun_out/tc64_sweep.txt
cat gpurun_out/l2_stream_probe.txt
