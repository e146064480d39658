#!/bin/bash
# Round-2 profiling pass (one GPU): ncu launch list of the driver-style bench command + one `--set full` capture per hot kernel.
# Numbers printed under ncu are never bench values; summaries are written by profiles/summarize.py into profiles/r02_*.md.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_b_ncu.log 2>&1
$NCU -k regex:dg_kronecker_march -s 10 -c 1 -o gpurun_out/r02_march python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cg --no-other-configs > /dev/null 2>&1
$NCU -k regex:dg_kronecker_mma -s 4 -c 1 -o gpurun_out/r02_mma_q3 python profiles/time_q3.py 133 > gpurun_out/r02_mma_q3.log 2>&1
B200FEM_Q3_SLAB=1 $NCU -k regex:dg_kronecker_slab -s 3 -c 1 -o gpurun_out/r02_slab_q3 python profiles/time_configs.py c5small > gpurun_out/r02_slab_q3.log 2>&1
$NCU -k regex:dg_kronecker_slab -s 3 -c 1 -o gpurun_out/r02_slab_q5 python profiles/time_configs.py c4 > gpurun_out/r02_slab_q5.log 2>&1
$NCU -k regex:lagrange_lattice -s 3 -c 1 -o gpurun_out/r02_lattice python profiles/time_lagrange.py > gpurun_out/r02_lattice.log 2>&1
$NCU -k regex:dg_quadrature -s 3 -c 1 -o gpurun_out/r02_quadrature python profiles/time_quadrature.py q2 > gpurun_out/r02_quadrature.log 2>&1
$NCU -k regex:lagrange_unstructured -s 215 -c 1 -o gpurun_out/r02_unstructured python profiles/time_unstructured.py > gpurun_out/r02_unstructured.log 2>&1
ls -la gpurun_out | grep r02_
