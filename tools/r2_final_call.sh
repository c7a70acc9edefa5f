#!/bin/bash
# Evidence bundle of round 2 (one GPU call): full GPU test suite, bench line, ncu launch list of the bench, ncu --set full of the
# tokeniser at both tile shapes, timelines, compute-sanitizer memcheck over every kernel family.
#   gpurun --timeout 2400 -- 'bash tools/r2_final_call.sh'
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 60 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list exit $?"
TCX_NCU=0 bash tools/tcx_bringup.sh > /dev/null 2>&1
for R in 64 96; do
  RQB200_TC_ROWS=$R timeout 300 ncu --set full --clock-control none --import-source on -k regex:rq_tcx -s 2 -c 1 -f -o gpurun_out/ncu_tcx_r$R tools/bin/tc_native_check 65536 768 3 3 /tmp/n.ids > gpurun_out/ncu_tcx_r$R.log 2>&1; echo "ncu r$R exit $?"
  RQB200_TC_ROWS=$R RQB200_TC_TRACE=1 timeout 60 tools/bin/tc_native_check 65536 768 3 20 /tmp/tl.ids > gpurun_out/timeline_tcx_r$R.txt 2>&1
  RQB200_TC_ROWS=$R timeout 60 tools/bin/tc_native_check 65536 768 3 50 /tmp/x.ids
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "sanitizer exit $?"; tail -4 gpurun_out/sanitizer_memcheck.log
