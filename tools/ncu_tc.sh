#!/bin/bash
# ncu captures of the tokeniser kernels through the Python-free harness (tools/tc_native_check.cu): one `--set full` capture
# per variant plus the memory-system counters that decide the round-2 question "is the kernel bound by L2 -> SM bytes?"
# (DESIGN.md 5.2b budget: codebook stream + x + Gram gathers against the L2 throughput cap).
#   usage (GPU box, repo root):  bash tools/ncu_tc.sh [variant ...]     variant = default | pair | fast | tma | tmapair | 64 | 64x4 | 64x8
# Outputs gpurun_out/ncu_tc_<variant>.ncu-rep and a CSV of the counters below; read them in the build container with
#   ncu -i gpurun_out/ncu_tc_<variant>.ncu-rep --page raw --csv
set -u
mkdir -p gpurun_out
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__data_pipe_lsu_wavefronts.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.sum,smsp__cycles_active.avg,lts__t_bytes.sum.pct_of_peak_sustained_elapsed,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed
for v in "${@:-default 64}"; do
  for variant in $v; do
    case $variant in
      default) envs="" ; pat="rq_tc_kernel" ;;
      pair)    envs="RQB200_TC_PAIR=1" ; pat="rq_tc_kernel" ;;
      fast)    envs="RQB200_TC_FASTSCAN=1" ; pat="rq_tc_kernel" ;;
      tma)     envs="RQB200_TC_TMA=1" ; pat="rq_tc_kernel" ;;
      tmapair) envs="RQB200_TC_TMA=1 RQB200_TC_PAIR=1" ; pat="rq_tc_kernel" ;;
      64)      envs="RQB200_TC_64=1" ; pat="rq_tc64_kernel" ;;
      64x4)    envs="RQB200_TC_64=4" ; pat="rq_tc64_kernel" ;;
      64x8)    envs="RQB200_TC_64=8" ; pat="rq_tc64_kernel" ;;
      *) echo "unknown variant $variant"; continue ;;
    esac
    env $envs timeout 300 ncu --metrics $METRICS --clock-control none -k regex:$pat -s 2 -c 1 --csv \
        --log-file gpurun_out/ncu_tc_${variant}_mem.csv tools/bin/tc_native_check 65536 768 3 2 > gpurun_out/ncu_tc_${variant}_mem.log 2>&1
    env $envs timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 -f \
        -o gpurun_out/ncu_tc_${variant} tools/bin/tc_native_check 65536 768 3 2 > gpurun_out/ncu_tc_${variant}.log 2>&1
  done
done
ls -la gpurun_out | tail -12
