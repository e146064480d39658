set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py 2>&1 | tail -5
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 1000 --warmup 20 --no-cg 2>&1 | tail -1 > gpurun_out/bench_n2_fused.json
B200FEM_NO_FUSED_SEND=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 1000 --warmup 20 --no-cg 2>&1 | tail -1 > gpurun_out/bench_n2_unfused.json
python - <<'PY'
import json
for f in ("fused","unfused"):
    try:
        d=json.load(open(f"gpurun_out/bench_n2_{f}.json")); print(f, d["ms_per_step"]*1e3, "us", d["value"]/1e9, "GDoF/s", d["config"]["kernel"], d.get("multi_gpu_diag"))
    except Exception as e: print(f, "failed", e)
PY
