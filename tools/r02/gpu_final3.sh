#!/bin/bash
mkdir -p gpurun_out
T=r02z
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"
TAMM_PARITY_REPORT=$PWD/gpurun_out/${T}_parity_report.jsonl python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_gpu_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02z_bench_full.json').read().strip().splitlines()[-1])
print('FULL value %.4g evals/s %.2f ms/step; stack frac %.3f (nominal %.3f); returns frac %.3f; traffic %s; e2e %.4g; e2e_power %.4g; cpu %s %.4g; launches %d' % (
    d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_of_nominal'], d['roofline_returns']['frac'], d['roofline']['traffic'],
    d['e2e']['value'], d['e2e_power']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['value'], d['gpu_launches']))
PY
