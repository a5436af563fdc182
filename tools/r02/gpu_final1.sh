#!/bin/bash
# round 2 final single-GPU record: parity suite (+ measured reports), smoke, default bench, every configuration, reference arm,
# launch list and ncu --set full of the two kernels of the step
mkdir -p gpurun_out
T=r02
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${T}_smoke.log
rm -f gpurun_out/${T}_parity_report.jsonl
TAMM_PARITY_REPORT=$PWD/gpurun_out/${T}_parity_report.jsonl python -m pytest tests -m gpu -q --durations=10 > gpurun_out/${T}_gpu_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${T}_gpu_tests.log; tail -4 gpurun_out/${T}_gpu_tests.log
python bench.py > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err; echo "bench rc=$?"
tools/bench_all_configs.sh $T
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; cut -c1-400 gpurun_out/${T}_bench_reference.json
CMD2="python bench.py --profiles 128 --batch 128 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD2 > gpurun_out/${T}_plain128.json 2> gpurun_out/${T}_plain128.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $CMD2 > gpurun_out/${T}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stack_kernel -s 3 -c 1 -o gpurun_out/${T}_stack $CMD2 > gpurun_out/${T}_ncu_stack.log 2>&1
ncu --set full --clock-control none -k regex:returns_fused -s 3 -c 1 -o gpurun_out/${T}_returns $CMD2 > gpurun_out/${T}_ncu_returns.log 2>&1
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_full.json').read().strip().splitlines()[-1])
print('FULL value %.4g evals/s %.2f ms/step; stack frac %.3f (nominal %.3f) launch %.3f ms; returns frac %.3f; e2e %.4g; e2e_power %.4g; cpu %s %.4g' % (
    d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_of_nominal'], d['roofline']['avg_launch_ms'], d['roofline_returns']['frac'],
    d['e2e']['value'], d['e2e_power']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['value']))
print('library md5', d['library']['md5'])
PY
du -sh gpurun_out
