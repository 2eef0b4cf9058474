P=29800
run() { N=$1; shift; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P "$@"; P=$((P+1)); }
timeout 200 python bench.py --steps 50 --warmup 10 --no-cpu > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err
run 2 bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/r2_scale_n2.json 2> gpurun_out/r2_scale_n2.err
run 4 bench.py --gpus 4 --steps 50 --warmup 10 > gpurun_out/r2_scale_n4.json 2> gpurun_out/r2_scale_n4.err
run 2 bench_cfg4.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_cfg4_n2.json 2> gpurun_out/r2_cfg4_n2.err
run 4 bench_cfg4.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_cfg4_n4.json 2> gpurun_out/r2_cfg4_n4.err
python - <<'PY'
import json
def load(f):
    for l in open(f):
        if l.startswith("{"): return json.loads(l)
for f in ("r2_scale_n1", "r2_scale_n2", "r2_scale_n4", "r2_cfg4_n2", "r2_cfg4_n4"):
    try:
        d = load(f"gpurun_out/{f}.json"); print(f, round(d["value"] / 1e9, 3), "G", round(d["ms_per_step"], 4), "ms e2e", round(d.get("e2e", {}).get("value", 0) / 1e9, 2), d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-800:])
PY
