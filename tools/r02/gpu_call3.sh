#!/bin/bash
# round 2, third GPU call (ONE GPU): kernels after the solver restructuring, graph-replayed single-stack configs, ncu captures
mkdir -p gpurun_out
T=r02c
CMD="python bench.py --profiles 128 --batch 128 --steps 5 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/${T}_plain128.json 2> gpurun_out/${T}_plain128.err
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_auto.json 2> gpurun_out/${T}_auto.err
python - <<'PY'
import json
for f in ('r02c_plain128', 'r02c_auto'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, '%.4g evals/s, %.2f ms/step, stack %.3f ms frac %.3f, returns %.3f ms frac %.3f, B=%s' % (d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline_returns']['avg_launch_ms'], d['roofline_returns']['frac'], d['config'].get('profiles_per_launch')))
    except Exception as e:
        print(f, 'failed', e)
PY
python -m pytest tests -m gpu -q -x > gpurun_out/${T}_gpu_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${T}_gpu_tests.log; tail -3 gpurun_out/${T}_gpu_tests.log
tools/bench_all_configs.sh $T
CMD2="python bench.py --profiles 128 --batch 128 --steps 2 --warmup 3 --no-e2e --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $CMD2 > gpurun_out/${T}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stack_kernel -s 3 -c 1 -o gpurun_out/${T}_stack $CMD2 > gpurun_out/${T}_ncu_stack.log 2>&1
ncu --set full --clock-control none -k regex:returns_fused -s 3 -c 1 -o gpurun_out/${T}_returns $CMD2 > gpurun_out/${T}_ncu_returns.log 2>&1
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
