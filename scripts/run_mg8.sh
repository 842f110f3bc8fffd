# final 8-GPU run: two plan variants of the library's multi-GPU path, then bench.py under torchrun with the faster one
pick() { python - "$@" <<'PY'
import json,sys
best=None
for spec in sys.argv[1:]:
    name,path=spec.split("=",1)
    try:
        r=json.loads(open(path).read().strip().splitlines()[-1]); ms=r["ms_device"]
    except Exception: continue
    if best is None or ms<best[1]: best=(name,ms)
print(best[0] if best else "default")
PY
}
timeout 200 python scripts/mg_bench.py lap7 128 8 3 0 2>&1 | tail -1 > gpurun_out/mg8_default.json; cat gpurun_out/mg8_default.json
SSB200_DIST_BALANCE=1 timeout 200 python scripts/mg_bench.py lap7 128 8 3 0 2>&1 | tail -1 > gpurun_out/mg8_balance.json; cat gpurun_out/mg8_balance.json
SSB200_MG_PULL_CTAS=148 timeout 200 python scripts/mg_bench.py lap7 128 8 3 0 2>&1 | tail -1 > gpurun_out/mg8_pull148.json; cat gpurun_out/mg8_pull148.json
B=$(pick default=gpurun_out/mg8_default.json balance=gpurun_out/mg8_balance.json pull148=gpurun_out/mg8_pull148.json)
echo "fastest variant: $B"
[ "$B" = balance ] && export SSB200_DIST_BALANCE=1
[ "$B" = pull148 ] && export SSB200_MG_PULL_CTAS=148
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -c 1800 gpurun_out/r2_bench_n8.json; tail -3 gpurun_out/r2_bench_n8.err
