#!/bin/bash
# Event timelines (CTA 0) of the tokeniser variants through the Python-free harness: the first thing to run in round 2.
#   bash tools/tc_timelines.sh            -> gpurun_out/timeline_<variant>.txt for default, tma, 64, 64x4
# Each file also carries the event-timed ms of the production instantiation (measured before the traced run).
mkdir -p gpurun_out
run() {  # name, env...
  local name=$1; shift
  env "$@" RQB200_TC_TRACE=1 timeout 60 tools/bin/tc_native_check 65536 768 3 20 /tmp/tl_$name.ids > gpurun_out/timeline_$name.txt 2>&1
  echo "$name: exit $? $(head -1 gpurun_out/timeline_$name.txt | cut -c1-200)"
}
run default RQB200_TC_64=0
run fast RQB200_TC_FASTSCAN=1
run tma RQB200_TC_TMA=1
run tmapair RQB200_TC_TMA=1 RQB200_TC_PAIR=1
run tmapairfast RQB200_TC_TMA=1 RQB200_TC_PAIR=1 RQB200_TC_FASTSCAN=1
run 64 RQB200_TC_64=1
run 64g2 RQB200_TC_64=1 RQB200_TC64_GROUPS=2
run 64x4 RQB200_TC_64=4
for v in fast tma tmapair tmapairfast 64 64g2 64x4; do cmp /tmp/tl_default.ids /tmp/tl_$v.ids && echo "IDS_IDENTICAL $v"; done
