F="--steps 200 --warmup 20 --no-cpu-baseline --no-cg --no-other-configs"
for i in 1 2; do
python bench.py $F 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('persist  ', d['ms_per_step'], d['linear_apply']['ms_per_step'])"
B200FEM_NO_L2_PERSIST=1 python bench.py $F 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nopersist', d['ms_per_step'], d['linear_apply']['ms_per_step'])"
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cg --no-other-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('persist short', d['ms_per_step'], d['linear_apply']['ms_per_step'])"
python -m pytest tests/test_gpu_march.py -x -q 2>&1 | tail -2
