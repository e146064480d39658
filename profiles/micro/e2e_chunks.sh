for c in 2 4 8 16; do B200FEM_PIPE_CHUNKS=$c python bench.py --steps 50 --warmup 5 --no-cg --no-other-configs --no-cpu-baseline --e2e-steps 30 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('chunks $c e2e', round(d['e2e']['value']/1e9,3), 'GDoF/s', round(7077888/d['e2e']['value']*1e3,3), 'ms')"; done
B200FEM_NO_PIPELINE=1 python bench.py --steps 50 --warmup 5 --no-cg --no-other-configs --no-cpu-baseline --e2e-steps 30 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('no pipeline e2e', round(d['e2e']['value']/1e9,3), 'GDoF/s')"
