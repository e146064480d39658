#!/bin/bash
# one-GPU validation pass: GPU tests, smoke, bench (driver-style short run and long run)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -5 gpurun_out/gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "bench short rc=$?"
python bench.py --no-cpu-baseline > gpurun_out/bench_long.json 2> gpurun_out/bench_long.err; echo "bench long rc=$?"
python - <<'PY'
import json
for f in ("bench_short", "bench_long"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "linear", d["linear_apply"]["ms_per_step"], d["linear_apply"]["frac"], "e2e", d["e2e"]["value"])
        for k, v in (d.get("configs") or {}).items():
            print("  ", k, {a: v[a] for a in v if a.endswith("_ms")}, v["roofline"]["frac"], v["roofline"].get("fp64", {}).get("frac"))
        for k, v in (d.get("cg") or {}).items():
            print("  ", k, v["s_per_iteration"], v["roofline"]["frac"])
        w = d.get("weak_scaling_c5")
        if w: print("   C5", w["affine_ms"], w["linear_ms"], w["roofline_linear"]["frac"])
        print("   cpu", d.get("cpu_baseline"))
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
tail -n 3 gpurun_out/bench_short.err; tail -n 3 gpurun_out/bench_long.err
