#!/bin/bash
# N-GPU tuning of the pipelined gather: NCCL CTA budget x block size (power and fields only)
N=${1:?ranks}; T=${2:?tag}; mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
i=0
for cfg in "8 0" "16 0" "16 16"; do
  set -- $cfg; i=$((i+1))
  NCCL_MAX_CTAS=$1 $RUN bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-cpu --gather power,fields --gather-mode nccl --gather-block $2 > gpurun_out/${T}_tune${i}_n$N.json 2> gpurun_out/${T}_tune${i}_n$N.err
done
python - "$N" "$T" <<'PY'
import json, sys
n, t = sys.argv[1], sys.argv[2]
for i in (1, 2, 3):
    try:
        d = json.loads(open('gpurun_out/%s_tune%d_n%s.json' % (t, i, n)).read().strip().splitlines()[-1])
        g = d['gather']
        print(i, g.get('nccl_env'), 'compute %.2f ms/step; block %d x %d' % (d['ms_per_step'], g['block_profiles'], g['blocks_per_rank']))
        for k in ('power', 'fields'):
            x = g[k]
            print('   %-6s back-to-back %.2f ms, pipelined nccl %.2f ms (floor %.2f) %s %s' % (k, x.get('ms', 0), x.get('pipelined_ms_per_step', 0), x.get('link_floor_ms', 0), x.get('pipelined_check'), x.get('error', '')))
    except Exception as e:
        print(i, 'failed', e)
PY
