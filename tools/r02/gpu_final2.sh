#!/bin/bash
# round 2: after the exact-zero fix of the fused returns -- deep-stack test, small configurations, alpha = 0 capture
mkdir -p gpurun_out
T=r02n
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "deep_stack or fixture or config0 or polarimetric" > gpurun_out/${T}_gpu_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${T}_gpu_tests.log; tail -3 gpurun_out/${T}_gpu_tests.log
for c in 0 1; do
  python bench.py --config $c --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_cfg${c}_$T.json 2> gpurun_out/bench_cfg${c}_$T.err
done
python bench.py --config 1 --alpha-deg 0 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_cfg1_a0_$T.json 2> gpurun_out/bench_cfg1_a0_$T.err
python bench.py --profiles 128 --batch 128 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_plain128.json 2> gpurun_out/${T}_plain128.err
python - <<'PY'
import json
for c in ('0', '1', '1_a0'):
    d = json.loads(open('gpurun_out/bench_cfg%s_r02n.json' % c).read().strip().splitlines()[-1])
    print('cfg%-5s %.4g evals/s %.4f ms/step (plain %.4f) stack %.4f ms, returns+scan %.4f ms, e2e %.4g' % (c, d['value'], d['ms_per_step'], d['launch_mode']['plain_launch_ms_per_step'], d['roofline']['avg_launch_ms'], d['returns_ms_per_step'], d['e2e']['value']))
d = json.loads(open('gpurun_out/r02n_plain128.json').read().strip().splitlines()[-1])
print('plain128 stack %.3f ms frac %.3f returns %.3f ms frac %.3f traffic %s' % (d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline_returns']['avg_launch_ms'], d['roofline_returns']['frac'], d['roofline']['traffic']))
PY
CMD3="python bench.py --config 1 --alpha-deg 0 --steps 3 --warmup 3 --no-e2e --no-cpu --no-graph"
ncu --set full --clock-control none -k regex:stack_kernel -s 3 -c 1 -o gpurun_out/${T}_stack_cfg1_a0 $CMD3 > gpurun_out/${T}_ncu_stack_a0.log 2>&1
CMD4="python bench.py --config 1 --steps 3 --warmup 3 --no-e2e --no-cpu --no-graph"
ncu --set full --clock-control none -k regex:stack_kernel -s 3 -c 1 -o gpurun_out/${T}_stack_cfg1_a10 $CMD4 > gpurun_out/${T}_ncu_stack_a10.log 2>&1
du -sh gpurun_out
