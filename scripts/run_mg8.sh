timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
txt=[l for l in open("gpurun_out/r2_bench_n8.json").read().splitlines() if l.startswith("{")]
r=json.loads(txt[-1])
print("value", r["value"], "ms", r["ms_per_step"], "e2e", r["e2e"]["value"], r["e2e"]["ms_per_step"], "solve", r["solve"]["ms"], r["solve"]["resid_2norm_rel"])
print("hbm", r["config"]["hbm_GB_per_gpu"], "nvlink", r["config"]["nvlink_GB_per_factorization"])
for e in r.get("extra_configs", []): print(e)
PY
tail -3 gpurun_out/r2_bench_n8.err
