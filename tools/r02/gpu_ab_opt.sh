#!/bin/bash
# A/B: in-tree library (candidate) against variants/libhead.so (the committed kernels), then the GPU suite on the candidate
mkdir -p gpurun_out
AB_TESTS=0 tools/ab_variant.sh variants/libhead.so r02opt
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02opt_auto.json 2> gpurun_out/r02opt_auto.err
TAMM_B200_LIB=$PWD/variants/libhead.so python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02opt_auto_head.json 2> gpurun_out/r02opt_auto_head.err
python - <<'PY'
import json
for f in ('r02opt_auto', 'r02opt_auto_head'):
    d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
    print(f, '%.4g evals/s %.2f ms/step stack frac %.3f returns frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline_returns']['frac']))
PY
python -m pytest tests -m gpu -x -q > gpurun_out/r02opt_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02opt_gpu_tests.log
