#!/bin/bash
# multi-GPU validation pass: parity check (peer-memory and NCCL transports) and bench.py at N = $1 ranks
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 tests/mgpu_check.py > gpurun_out/mgpu_check_n$N.log 2>&1; echo "mgpu_check rc=$?"; grep -E "mgpu_check OK|Error|error|assert" gpurun_out/mgpu_check_n$N.log | head -20
if [ "$2" != "quick" ]; then
B200FEM_NO_P2P=1 timeout 900 $TR --master-port 29512 tests/mgpu_check.py > gpurun_out/mgpu_check_nccl_n$N.log 2>&1; echo "mgpu_check (nccl) rc=$?"; grep -E "mgpu_check OK|Error|error|assert" gpurun_out/mgpu_check_nccl_n$N.log | head -20
fi
timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_short.json 2> gpurun_out/bench_n${N}_short.err; echo "bench short rc=$?"
timeout 900 $TR --master-port 29514 bench.py --gpus $N --steps 1000 --warmup 20 --no-other-configs --no-cg --no-parity > gpurun_out/bench_n${N}_long.json 2> gpurun_out/bench_n${N}_long.err; echo "bench long rc=$?"
python - $N <<'PY'
import json, sys
N = sys.argv[1]
for f in (f"bench_n{N}_short", f"bench_n{N}_long"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "us/step", 1e3 * d["ms_per_step"], "GDoF/s", d["value"] / 1e9, "linear us", 1e3 * d["linear_apply"]["ms_per_step"], "e2e", d["e2e"]["value"] / 1e9, "host us", d["host_issue_us_per_step"])
        print("   diag", d.get("multi_gpu_diag"))
        for k, v in (d.get("cg") or {}).items():
            print("   cg", k, v["s_per_iteration"], v["schedule"])
        w = d.get("weak_scaling_c5")
        if w: print("   C5 affine ms", w["affine_ms"], "linear ms", w["linear_ms"], "GDoF/s", w["value"] / 1e9)
        print("   parity", d.get("parity"))
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
tail -n 5 gpurun_out/bench_n${N}_short.err
