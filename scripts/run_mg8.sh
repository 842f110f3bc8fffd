set -x
export SSB200_MG_TRACE=1
timeout 200 python scripts/mg_bench.py lap7 128 8 3 0 2>&1 | tail -3
cp gpurun_out/mg_trace_lap7128_n8.npz gpurun_out/mg_trace_n8_keep.npz
unset SSB200_MG_TRACE
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -c 2500 gpurun_out/r2_bench_n8.json; tail -5 gpurun_out/r2_bench_n8.err
