#!/bin/bash
# round 2, first GPU call: parity suite, smoke, default bench, every configuration, reference arm, variant A/B
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r02_smoke.log
python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r02_gpu_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r02_gpu_tests.log
tail -30 gpurun_out/r02_gpu_tests.log
python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err; echo "bench rc=$?"
cat gpurun_out/r02_bench_full.json
tools/bench_all_configs.sh r02a
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cat gpurun_out/r02_bench_reference.json
# variants: 192 x 2 CTAs on the big configuration, 384 x 1 on the single-profile one
for v in 5 6; do
  TAMM_STACK_VARIANT=$v python bench.py --profiles 128 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_var${v}_cfg4.json 2> gpurun_out/r02_var${v}_cfg4.err
  TAMM_STACK_VARIANT=$v python bench.py --config 1 --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_var${v}_cfg1.json 2> gpurun_out/r02_var${v}_cfg1.err
done
python - <<'PY'
import json
for f in ('r02_var5_cfg4', 'r02_var6_cfg4', 'r02_var5_cfg1', 'r02_var6_cfg1'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, '%.4g evals/s, stack %.3f ms, frac %.3f' % (d['value'], d['roofline']['avg_launch_ms'], d['roofline']['frac']))
    except Exception as e:
        print(f, 'failed', e)
PY
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_smi.csv
lscpu | head -20 > gpurun_out/r02_host.txt; free -g >> gpurun_out/r02_host.txt; nvidia-smi topo -m >> gpurun_out/r02_host.txt 2>&1
