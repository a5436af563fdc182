#!/bin/bash
# One GPU call for the direct-returns path (tamm_stack_returns), sized for the last ~3 GPU-minutes of the round:
#   1. tests/test_zz_gpu_direct.py            direct path == two-call path on small shapes (all F regimes, cold path, ABI, ensemble)
#   2. bench configs[4], direct on / off   10 timed steps each, no e2e, no CPU arm
#   3. the full-size unchunked launch test with the direct path on (oracle + arbiter on three profiles)
# Everything lands in gpurun_out/r02d_*; each step has its own timeout so a late step cannot lose the earlier results.
mkdir -p gpurun_out
T0=$SECONDS
timeout 60 python -m pytest tests/test_zz_gpu_direct.py -m gpu -x -q --runxfail > gpurun_out/r02d_tests_direct.log 2>&1
echo "rc=$? t=$((SECONDS-T0))s" >> gpurun_out/r02d_tests_direct.log
tail -3 gpurun_out/r02d_tests_direct.log
B="python bench.py --config 4 --steps 10 --warmup 3 --no-e2e --no-cpu"
TAMM_B200_DIRECT_RETURNS=1 timeout 70 $B > gpurun_out/r02d_bench_direct_on.json 2> gpurun_out/r02d_bench_direct_on.err
echo "on rc=$? t=$((SECONDS-T0))s"
TAMM_B200_DIRECT_RETURNS=0 timeout 70 $B > gpurun_out/r02d_bench_direct_off.json 2> gpurun_out/r02d_bench_direct_off.err
echo "off rc=$? t=$((SECONDS-T0))s"
python - <<'PY'
import json
for k in ('on', 'off'):
    try:
        d = json.loads(open('gpurun_out/r02d_bench_direct_%s.json' % k).read().strip().splitlines()[-1])
        r, q = d['roofline'], d.get('roofline_returns') or {}
        print(k, 'value %.4g ms/step %.2f | %s %.3f ms frac %.3f | %s %.3f ms frac %.3f' % (
            d['value'], d['ms_per_step'], r['kernel'][:22], r['avg_launch_ms'], r['frac'], q.get('kernel'), q.get('avg_launch_ms', float('nan')),
            q.get('frac', float('nan'))))
    except Exception as e:
        print(k, 'failed:', repr(e))
PY
TAMM_B200_DIRECT_RETURNS=1 TAMM_PARITY_REPORT=gpurun_out/r02d_parity_report.jsonl timeout 80 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k unchunked > gpurun_out/r02d_tests_unchunked.log 2>&1
echo "rc=$? t=$((SECONDS-T0))s" >> gpurun_out/r02d_tests_unchunked.log
tail -2 gpurun_out/r02d_tests_unchunked.log
