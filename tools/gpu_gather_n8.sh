#!/bin/bash
N=${1:?ranks}; T=${2:?tag}; mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523"
$RUN bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-cpu --gather-mode both > gpurun_out/${T}_a_n$N.json 2> gpurun_out/${T}_a_n$N.err
$RUN bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-cpu --gather power,fields --gather-mode peer --push-ctas 4 > gpurun_out/${T}_b_n$N.json 2> gpurun_out/${T}_b_n$N.err
python - "$N" "$T" <<'PY'
import json, sys
n, t = sys.argv[1], sys.argv[2]
for tag in ('a', 'b'):
    try:
        d = json.loads(open('gpurun_out/%s_%s_n%s.json' % (t, tag, n)).read().strip().splitlines()[-1])
        g = d['gather']
        print(tag, 'compute %.2f ms/step; schedule %s' % (d['ms_per_step'], g.get('block_schedule')))
        for k in ('power', 'fields', 'jones'):
            if k not in g: continue
            x = g[k]
            pip = {kk.replace('pipelined_', '').replace('_ms_per_step', ''): round(v, 2) for kk, v in x.items() if kk.endswith('_ms_per_step') and kk.startswith('pipelined')}
            chk = {kk: v for kk, v in x.items() if kk.endswith('_check') or kk.endswith('_error')}
            print('   %-6s back-to-back %.2f ms, floor %.2f; pipelined step (ms): %s  %s' % (k, x.get('ms', 0), x.get('link_floor_ms', 0), pip, chk))
    except Exception as e:
        print(tag, 'failed', e)
PY
tail -3 gpurun_out/${T}_a_n$N.err
