# bench.py under torchrun on N GPUs of one box:  gpurun --gpus N -- 'bash profiles/microbench/scale_n.sh N'
N=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 50 --warmup 10 > gpurun_out/scale_n$N.out 2> gpurun_out/scale_n$N.err
grep "^{" gpurun_out/scale_n$N.out > gpurun_out/scale_n$N.json
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/scale_n$N.json")); print($N, d["value"], d["ms_per_step"], d["e2e"]["value"], d["details"].get("gradient_exchange"))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/scale_n$N.err").read()[-2000:])
PY
