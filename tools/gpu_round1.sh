#!/usr/bin/env bash
# One gpurun call: tests, both bench arms, ncu launch list and one full capture of the top kernel.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r01.json 2> gpurun_out/bench_ref_r01.err; cat gpurun_out/bench_ref_r01.json
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -3 gpurun_out/bench_r01.err
cat gpurun_out/bench_r01.json
SHORT="python bench.py --steps 2 --warmup 3 --frames 2368 --e2e-frames 592 --no-cpu-baseline"
$SHORT > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv $SHORT > gpurun_out/ncu1.log 2>&1
$SHORT > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ldpc_layered_i8 -s 3 -c 2 -f -o gpurun_out/prof_layered_r01 $SHORT > gpurun_out/ncu2.log 2>&1
python tools/nb_bench.py > gpurun_out/nb_bench.txt 2>&1; cat gpurun_out/nb_bench.txt
ls -la gpurun_out
