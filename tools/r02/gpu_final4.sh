#!/bin/bash
# round 2: ncu capture and default bench line of the final kernels (after the integer-pipe tests of the stack kernel)
mkdir -p gpurun_out
T=r02y
CMD2="python bench.py --profiles 128 --batch 128 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD2 > gpurun_out/${T}_plain128.json 2> gpurun_out/${T}_plain128.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $CMD2 > gpurun_out/${T}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stack_kernel -s 3 -c 1 -o gpurun_out/${T}_stack $CMD2 > gpurun_out/${T}_ncu_stack.log 2>&1
ncu --set full --clock-control none -k regex:returns_fused -s 3 -c 1 -o gpurun_out/${T}_returns $CMD2 > gpurun_out/${T}_ncu_returns.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 --cpu-seconds 6 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference arm rc=$?"; cut -c1-300 gpurun_out/${T}_bench_reference.json
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02y_bench_full.json').read().strip().splitlines()[-1])
print('FULL value %.4g evals/s %.2f ms/step; stack frac %.3f (nominal %.3f) launch %.3f ms; returns frac %.3f; traffic %s; e2e %.4g; e2e_power %.4g; cpu %s %.4g' % (
    d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_of_nominal'], d['roofline']['avg_launch_ms'], d['roofline_returns']['frac'], d['roofline']['traffic'],
    d['e2e']['value'], d['e2e_power']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['value']))
PY
du -sh gpurun_out
