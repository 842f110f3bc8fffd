O=gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > $O/r2_gpu_tests_final.log 2>&1; cat $O/r2_gpu_tests_final.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; tail -c 600 $O/r2_bench_n1.json; tail -2 $O/r2_bench_n1.err
timeout 120 python scripts/fp64_peak.py > $O/r2_fp64_peak.json 2>&1; cat $O/r2_fp64_peak.json
export SSB200_LOOKAHEAD=0
cap() { local name=$1 skip=$2; timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_sub_kernel -s $skip -c 1 -f -o $O/r2_$name python scripts/profile_step.py lap7 128 1 > $O/r2_cap_$name.log 2>&1; }
cap gemm128_k1024 533
cap gemm128_update 515
cap gemm128_k64 14
cap gemm64 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:potrf_block_kernel4 -s 700 -c 1 -f -o $O/r2_potrf4 python scripts/profile_step.py lap7 128 1 > $O/r2_cap_potrf4.log 2>&1
ls -la $O/r2_gemm*.ncu-rep $O/r2_potrf4.ncu-rep
