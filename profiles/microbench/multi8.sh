P=29700
run() { N=$1; shift; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P "$@"; P=$((P+1)); }
CUDA_VISIBLE_DEVICES=0,1 timeout 300 python -m pytest tests/test_dataparallel_gpu.py -m gpu -q -x > gpurun_out/r2_dp2_test.log 2>&1; tail -3 gpurun_out/r2_dp2_test.log
run 8 bench.py --gpus 8 --steps 50 --warmup 10 > gpurun_out/r2_scale_n8.json 2> gpurun_out/r2_scale_n8.err
run 8 bench.py --gpus 8 --steps 50 --warmup 10 --nccl-allreduce > gpurun_out/r2_scale_n8_nccl.json 2> gpurun_out/r2_scale_n8_nccl.err
run 8 bench_cfg4.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_cfg4_n8.json 2> gpurun_out/r2_cfg4_n8.err
run 8 bench_configs.py --dp > gpurun_out/r2_cfg5_dp8.jsonl 2> gpurun_out/r2_cfg5_dp8.err
python - <<'PY'
import json
for f in ("r2_scale_n8", "r2_scale_n8_nccl", "r2_cfg4_n8"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], d.get("e2e", {}).get("value"), d.get("details", {}).get("gradient_exchange"), d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
print(open("gpurun_out/r2_cfg5_dp8.jsonl").read()[:600])
PY
