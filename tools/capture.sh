#!/bin/bash
# Standard measurement sequence for one build (run on the GPU box through gpurun, ONE GPU):
#   tools/capture.sh TAG        e.g.  gpurun --timeout 900 -- 'tools/capture.sh r02_v16'
# 1. short bench without ncu (must exit 0 first), 2. launch list (cold, serialised: shares of a step),
# 3. ncu --set full of the stack kernel and of the fused returns kernel, 4. the full default bench line.
# Everything lands in gpurun_out/ (keep a call's output below 64 MiB: two reports with --import-source are ~40 MiB);
# summarise the reports here with tools/ncu_summary.py and copy what is to be judged into profiles/.
T=${1:?tag}
mkdir -p gpurun_out
CMD="python bench.py --profiles 128 --batch 128 --steps 2 --warmup 3 --no-e2e --no-cpu"   # one chunked 128-profile launch per step
$CMD > gpurun_out/plain_$T.json 2> gpurun_out/plain_$T.err || { echo "plain run failed"; tail -5 gpurun_out/plain_$T.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_list_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stack_kernel -s 3 -c 1 -o gpurun_out/stack_$T $CMD > gpurun_out/ncu_stack_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:returns_fused -s 3 -c 1 -o gpurun_out/returns_$T $CMD > gpurun_out/ncu_returns_$T.log 2>&1
python bench.py > gpurun_out/bench_full_$T.json 2> gpurun_out/bench_full_$T.err
cat gpurun_out/plain_$T.json; echo; cat gpurun_out/bench_full_$T.json
