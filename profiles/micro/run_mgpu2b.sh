N=2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py 2>&1 | grep -E "mgpu_check|Error|error" | head -5
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 1000 --warmup 20 2>/dev/null | tail -1 > gpurun_out/bench_n2_pdl.json
B200FEM_NO_PDL=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 1000 --warmup 20 2>/dev/null | tail -1 > gpurun_out/bench_n2_nopdl.json
python - <<PY
import json
for f in ("pdl","nopdl"):
    try:
        d=json.load(open(f"gpurun_out/bench_n2_{f}.json")); print("N=2", f, round(d["ms_per_step"]*1e3,2), "us", round(d["value"]/1e9,1), "GDoF/s", d.get("multi_gpu_diag"), d["host_issue_us_per_step"])
    except Exception as e: print(f, "failed", e)
PY
