#!/bin/bash
mkdir -p gpurun_out
T=r02m
python bench.py --profiles 128 --batch 128 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_plain128.json 2> gpurun_out/${T}_plain128.err
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_auto.json 2> gpurun_out/${T}_auto.err
python - <<'PY'
import json
for f in ('r02m_plain128', 'r02m_auto'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, '%.4g evals/s, %.2f ms/step, stack %.3f ms frac %.3f, returns %.3f ms frac %.3f, B=%s' % (d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline_returns']['avg_launch_ms'], d['roofline_returns']['frac'], d['config'].get('profiles_per_launch')))
    except Exception as e:
        print(f, 'failed', e)
PY
TAMM_PARITY_REPORT=$PWD/gpurun_out/${T}_parity_report.jsonl python -m pytest tests -m gpu -q -x > gpurun_out/${T}_gpu_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${T}_gpu_tests.log; tail -3 gpurun_out/${T}_gpu_tests.log
