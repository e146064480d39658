F="--steps 20 --warmup 5 --no-cpu-baseline --no-cg --no-other-configs"
for i in 1 2 3; do python bench.py $F 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default', d['ms_per_step'], d['linear_apply']['ms_per_step'], d['clocks'])"; done
for i in 1 2 3; do B200FEM_BENCH_NOSAMPLER=1 python bench.py $F 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nosampler', d['ms_per_step'], d['linear_apply']['ms_per_step'])"; done
B200FEM_BENCH_STEPTIMES=1 python bench.py $F 2>&1 >/dev/null | grep "step times"
python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-cg --no-other-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('200 steps', d['ms_per_step'], d['linear_apply']['ms_per_step'])"
