#!/bin/bash
# Multi-GPU measurements on ONE box:  gpurun --gpus N --timeout 1500 -- 'tools/gpu_multi.sh N TAG [tests]'
# config 4 at N ranks (value + e2e + the three gather sizes, pipelined), then the sharded single-stack configs.
N=${1:?ranks}; T=${2:?tag}; mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
if [ "$3" = "tests" ]; then
  python -m pytest tests/test_gpu_sharded.py -q > gpurun_out/${T}_sharded_tests.log 2>&1; echo "sharded tests rc=$?" | tee -a gpurun_out/${T}_sharded_tests.log; tail -3 gpurun_out/${T}_sharded_tests.log
fi
$RUN bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${T}_cfg4_n$N.json 2> gpurun_out/${T}_cfg4_n$N.err; echo "cfg4 rc=$?"
for c in 1 2 3; do
  $RUN bench.py --gpus $N --config $c --steps 10 --warmup 3 > gpurun_out/${T}_cfg${c}_n$N.json 2> gpurun_out/${T}_cfg${c}_n$N.err; echo "cfg$c rc=$?"
done
nvidia-smi topo -m > gpurun_out/${T}_topo_n$N.txt 2>&1
python - "$N" "$T" <<'PY'
import json, sys
n, t = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open('gpurun_out/%s_cfg4_n%s.json' % (t, n)).read().strip().splitlines()[-1])
    print('cfg4 N=%s value %.4g evals/s (%.2f ms/step)' % (n, d['value'], d['ms_per_step']))
    e = d.get('e2e', {}); p = d.get('e2e_power', {})
    print('  e2e %.4g evals/s (%.1f ms/call, %.1f GB/s/rank; bare d2h %.1f GB/s/rank, %.1f all ranks)   e2e_power %.4g' % (
        e.get('value', 0), e.get('ms_per_call', 0), e.get('d2h_gbs_per_rank', 0), e.get('bare_d2h_gbs_per_rank', 0), e.get('bare_d2h_gbs_all_ranks', 0), p.get('value', 0)))
    g = d.get('gather', {})
    for k in ('power', 'fields', 'jones'):
        if k in g:
            x = g[k]
            print('  gather %-6s: back-to-back %.2f ms (algbw %.0f GB/s), pipelined step nccl %.2f ms / ce %s ms -> %.4g evals/s, link floor %.2f ms, checks %s %s %s %s' % (
                k, x.get('ms', 0), x.get('algbw_gbs', 0), x.get('pipelined_ms_per_step', 0), x.get('pipelined_ce_ms_per_step'), x.get('value_with_gather', 0), x.get('link_floor_ms', 0), x.get('pipelined_check'), x.get('pipelined_ce_check'), x.get('pipelined_ce_error', ''), x.get('error', '')))
except Exception as ex:
    print('cfg4 failed:', ex)
for c in (1, 2, 3):
    try:
        d = json.loads(open('gpurun_out/%s_cfg%d_n%s.json' % (t, c, n)).read().strip().splitlines()[-1])
        e = d.get('e2e') or {}
        print('cfg%d N=%s value %.4g evals/s (%.3f ms/step)  e2e %s' % (c, n, d['value'], d['ms_per_step'], e.get('value') or e.get('error')))
    except Exception as ex:
        print('cfg', c, 'failed:', ex)
PY
tail -3 gpurun_out/${T}_cfg4_n$N.err
