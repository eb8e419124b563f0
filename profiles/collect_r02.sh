#!/bin/bash
# Round-2 evidence run on ONE B200, from the repo root (everything lands under gpurun_out/):
#   GPU test suite, bench.py (own arm + reference arm), ncu launch list, ncu --set full captures of the three fused kernels,
#   end-of-training accuracy.  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
NCU="ncu --set full --import-source on --clock-control none -k regex:fused_loss_grad -c 1"
$NCU -s 3 -o gpurun_out/burgers python profiles/run_fused.py 100000 6 > gpurun_out/ncu_burgers.log 2>&1
$NCU -s 3 -o gpurun_out/nls python profiles/run_nls.py 20000 5 > gpurun_out/ncu_nls.log 2>&1
$NCU -s 4 -o gpurun_out/gen40 python profiles/time_generic.py 8x40 > gpurun_out/ncu_gen40.log 2>&1
python profiles/run_accuracy.py > gpurun_out/accuracy.json 2> gpurun_out/accuracy.err
ls -la gpurun_out
