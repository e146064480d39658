fmt='
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    for k,v in d.items(): print(k, "affine %.1f us %.1f GDoF/s | linear %.1f us %.1f GDoF/s %.1f TF" % (v["affine"]["ms"]*1e3, v["affine"]["dofs_per_s"]/1e9, v["linear"]["ms"]*1e3, v["linear"]["dofs_per_s"]/1e9, v["kron_tflops_linear"]))
'
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "slab or mol or pipeline" 2>&1 | tail -1
for wv in 0 1 2 4; do echo "== waves $wv"; B200FEM_SLAB_WAVES=$wv timeout 200 python profiles/time_configs.py c4 c5small 2>&1 | python -c "$fmt"; done
