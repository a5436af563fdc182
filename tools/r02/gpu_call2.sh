#!/bin/bash
# round 2, second GPU call: failed tests again + device solver test, configurations, returns-kernel A/B, ncu captures
mkdir -p gpurun_out
T=r02b
python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/${T}_gpu_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${T}_gpu_tests.log
tail -15 gpurun_out/${T}_gpu_tests.log
tools/bench_all_configs.sh $T
CMD="python bench.py --profiles 128 --steps 5 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/${T}_plain.json 2> gpurun_out/${T}_plain.err
TAMM_RETURNS_V16=1 $CMD > gpurun_out/${T}_v16.json 2> gpurun_out/${T}_v16.err
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --batch 313 > gpurun_out/${T}_b313.json 2> gpurun_out/${T}_b313.err
python - <<'PY'
import json
for f in ('r02b_plain', 'r02b_v16', 'r02b_b313'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, '%.4g evals/s, stack %.3f ms frac %.3f, returns %.3f ms frac %.3f' % (d['value'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline_returns']['avg_launch_ms'], d['roofline_returns']['frac']))
    except Exception as e:
        print(f, 'failed', e)
PY
CMD2="python bench.py --profiles 128 --steps 2 --warmup 3 --no-e2e --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $CMD2 > gpurun_out/${T}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stack_kernel -s 3 -c 1 -o gpurun_out/${T}_stack $CMD2 > gpurun_out/${T}_ncu_stack.log 2>&1
ncu --set full --clock-control none -k regex:returns_fused -s 3 -c 1 -o gpurun_out/${T}_returns $CMD2 > gpurun_out/${T}_ncu_returns.log 2>&1
TAMM_RETURNS_V16=1 ncu --set full --clock-control none -k regex:returns_fused16 -s 3 -c 1 -o gpurun_out/${T}_returns16 $CMD2 > gpurun_out/${T}_ncu_returns16.log 2>&1
# alpha = 0: the closed-form eigen path (configs[0], configs[1] at normal incidence)
CMD3="python bench.py --config 1 --alpha-deg 0 --steps 3 --warmup 3 --no-e2e --no-cpu"
ncu --set full --clock-control none -k regex:stack_kernel -s 3 -c 1 -o gpurun_out/${T}_stack_cfg1_a0 $CMD3 > gpurun_out/${T}_ncu_stack_a0.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8
