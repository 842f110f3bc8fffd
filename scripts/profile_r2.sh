#!/bin/bash
# Round-2 evidence run (one B200): launch list of the bench command, `ncu --set full` captures of every kernel, sanitizer logs.
# Outputs under gpurun_out/ ; scripts/summarize_profiles.py turns them into the text files committed under profiles/.
export SSB200_LOOKAHEAD=0 SSB200_SOLVE_GRAPH=0
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/r2_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $O/r2_bench_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_lap7_64.csv python scripts/profile_step.py lap7 64 1 > $O/r2_profile_step_64.log 2>&1
cap() { # name regex skip count workload...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o $O/r2_$name "$@" > $O/r2_cap_$name.log 2>&1
}
cap potrf3 potrf_block_kernel3 90 1 python scripts/profile_step.py lap7 64 1
cap trsm_tc trsm_tc_kernel 90 1 python scripts/profile_step.py lap7 64 1
cap trsm_rows trsm_rows_kernel 10 1 python scripts/profile_step.py lap7 64 1
cap gemm64 'gemm_nt_sub_kernel<64' 5 1 python scripts/profile_step.py lap7 64 1
cap scatter_relmap 'scatter_A_kernel|relmap_kernel|fill_int_kernel|factor_diag' 0 3 python scripts/profile_step.py lap7 64 1
cap lsolve_diag lsolve_diag_kernel 200 1 python scripts/profile_step.py lap7 64 1
cap lsolve_update lsolve_update_kernel 200 1 python scripts/profile_step.py lap7 64 1
cap ltsolve_update ltsolve_update_kernel 20 1 python scripts/profile_step.py lap7 64 1
cap ltsolve_diag ltsolve_diag_kernel 20 1 python scripts/profile_step.py lap7 64 1
SSB200_SOLVE_BLK=1 cap solve_blk solve_blk_kernel 2 2 python scripts/profile_step.py lap7 64 1
# the dominant kernel at the benchmark size: a K = 1024 trailing update of the root and a large-K descendant update
cap gemm128_k1024 'gemm_nt_sub_kernel<128' 700 1 python scripts/profile_step.py lap7 128 1
cap gemm128_update 'gemm_nt_sub_kernel<128' 14 1 python scripts/profile_step.py lap7 128 1
# sanitizers on small problems through the plain and the drop-in layer (both solve schedules)
timeout 900 compute-sanitizer --tool memcheck python scripts/gpu_first.py lap7:8 lap7:20 lap27:10 elas:5 > $O/r2_sanitizer_memcheck.log 2>&1
SSB200_SOLVE_BLK=1 SSB200_SOLVE_BLK_MIN=100 timeout 900 compute-sanitizer --tool memcheck python scripts/gpu_first.py lap7:20 > $O/r2_sanitizer_memcheck_blk.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python scripts/gpu_first.py lap7:12 > $O/r2_sanitizer_racecheck.log 2>&1
tail -3 $O/r2_sanitizer_*.log
ls -la $O/*.ncu-rep | wc -l
