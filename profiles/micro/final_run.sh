set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
python __graft_entry__.py smoke 2>&1 | tail -5
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; head -c 600 gpurun_out/bench.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_reference.json | head -c 400; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dg_kronecker_march -s 10 -c 1 -f -o gpurun_out/march python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cg --no-other-configs > /dev/null 2>&1
ls -la gpurun_out
