timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 5 --warmup 3 --no-extra > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
python - <<'PY'
import json
txt=[l for l in open("gpurun_out/r2_bench_n4.json").read().splitlines() if l.startswith("{")]
r=json.loads(txt[-1])
print("value", r["value"], "ms", r["ms_per_step"], "e2e", r["e2e"]["value"], r["e2e"]["ms_per_step"], "solve", r["solve"]["ms"], r["solve"]["resid_2norm_rel"], "hbm", r["config"]["hbm_GB_per_gpu"], "nvlink", r["config"]["nvlink_GB_per_factorization"])
PY
