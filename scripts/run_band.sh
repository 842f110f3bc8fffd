O=gpurun_out
(timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or oracle or reference_library or full_size" 2>&1 | tail -3)
EXP_TAG=band timeout 200 python scripts/exp_factor.py lap7 128 3 2>&1 | tail -1
EXP_TAG=band_serial SSB200_LOOKAHEAD=0 timeout 200 python scripts/exp_factor.py lap7 128 2 2>&1 | tail -1
SSB200_LOOKAHEAD=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_sub_kernel -s 515 -c 1 -f -o $O/r2_gemm128_update_band python scripts/profile_step.py lap7 128 1 > $O/r2_cap_gemm128_update_band.log 2>&1
ls -la $O/r2_gemm128_update_band.ncu-rep
