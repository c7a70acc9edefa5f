#!/bin/bash
# Final evidence bundle of round 2 (one GPU call, ~6 minutes): GPU test suite, bench line, ncu launch list of the bench, ncu --set
# full of the split GEMM / splitter / fused chain / SGEMM / backward / Gumbel row kernels, pipeline timings, sanitizer.
#   gpurun --timeout 1500 -- 'bash tools/r2_evidence.sh'
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 2500 gpurun_out/bench_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 60 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list exit $?"
python tools/real_pipeline.py > gpurun_out/real_pipeline.txt 2>&1; tail -2 gpurun_out/real_pipeline.txt
python tools/gemm_split_time.py > gpurun_out/gemm_split_time.txt 2>&1; cat gpurun_out/gemm_split_time.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gs_gemm|gs_split|gs_colmax|gs_reduce|rq_fused|rq_replay|sgemm_kernel|rq_bwd|gumbel|dist_finish|row_finish|sid_' -c 80 -f -o /tmp/ncu_kernels python tools/kernels_prof.py > gpurun_out/ncu_kernels.log 2>&1; echo "ncu kernels exit $?"
# the report itself stays on the box (gpurun_out/ is capped at 64 MiB): bring back the raw metric table
ncu -i /tmp/ncu_kernels.ncu-rep --page raw --csv > gpurun_out/ncu_kernels_raw.csv 2>/dev/null; wc -c gpurun_out/ncu_kernels_raw.csv
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "sanitizer exit $?"; tail -4 gpurun_out/sanitizer_memcheck.log
