#!/bin/bash
# gather-overlap experiment on N GPUs: high-priority NCCL stream on / off, block count
N=${1:?ranks}; T=${2:?tag}; mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
if [ "$3" = "tests" ]; then python -m pytest tests/test_gpu_sharded.py -q -k nccl > gpurun_out/${T}_sharded_tests.log 2>&1; echo "sharded tests rc=$?"; tail -5 gpurun_out/${T}_sharded_tests.log; fi
$RUN bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-cpu --gather-mode peer $EXTRA > gpurun_out/${T}_hp1_n$N.json 2> gpurun_out/${T}_hp1_n$N.err
for ct in 4 16; do $RUN bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-cpu --gather power,fields --gather-mode peer --push-ctas $ct > gpurun_out/${T}_ctas${ct}_n$N.json 2> gpurun_out/${T}_ctas${ct}_n$N.err; done
python - "$N" "$T" <<'PY'
import json, sys
n, t = sys.argv[1], sys.argv[2]
for tag in ('hp1', 'ctas4', 'ctas16'):
    try:
        d = json.loads(open('gpurun_out/%s_%s_n%s.json' % (t, tag, n)).read().strip().splitlines()[-1])
        g = d['gather']
        print(tag, 'compute %.2f ms/step; block %d x %d' % (d['ms_per_step'], g['block_profiles'], g['blocks_per_rank']))
        for k in ('power', 'fields', 'jones'):
            if k not in g: continue
            x = g[k]
            pip = {kk.replace('pipelined_', '').replace('_ms_per_step', ''): round(v, 2) for kk, v in x.items() if kk.endswith('_ms_per_step') and kk.startswith('pipelined')}
            chk = {kk: v for kk, v in x.items() if kk.endswith('_check') or kk.endswith('_error')}
            print('   %-6s back-to-back %.2f ms, floor %.2f; pipelined step (ms): %s  %s -> %s' % (k, x.get('ms', 0), x.get('link_floor_ms', 0), pip, chk, x.get('pipelined_transport')))
    except Exception as e:
        print(tag, 'failed', e)
PY
tail -5 gpurun_out/${T}_hp1_n$N.err
