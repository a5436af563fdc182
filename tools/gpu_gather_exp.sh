#!/bin/bash
# gather-overlap experiment on N GPUs: high-priority NCCL stream on / off, block count
N=${1:?ranks}; T=${2:?tag}; mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
for hp in 1 0; do
  TAMM_NCCL_HIGH_PRIORITY=$hp $RUN bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_hp${hp}_n$N.json 2> gpurun_out/${T}_hp${hp}_n$N.err
done
TAMM_NCCL_HIGH_PRIORITY=1 $RUN bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-cpu --gather-block 32 > gpurun_out/${T}_hp1_b32_n$N.json 2> gpurun_out/${T}_hp1_b32_n$N.err
TAMM_NCCL_HIGH_PRIORITY=1 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL $RUN bench.py --gpus $N --steps 2 --warmup 3 --no-e2e --no-cpu --gather fields > gpurun_out/${T}_dbg_n$N.json 2> gpurun_out/${T}_dbg_n$N.err
grep -i "nvls\|algo\|proto\|channels\|CE " gpurun_out/${T}_dbg_n$N.err | sort | uniq -c | sort -rn | head -20 > gpurun_out/${T}_nccl_info_n$N.txt
python - "$N" "$T" <<'PY'
import json, sys
n, t = sys.argv[1], sys.argv[2]
for tag in ('hp1', 'hp0', 'hp1_b32'):
    try:
        d = json.loads(open('gpurun_out/%s_%s_n%s.json' % (t, tag, n)).read().strip().splitlines()[-1])
        g = d['gather']
        print(tag, 'compute %.2f ms/step; block %d x %d' % (d['ms_per_step'], g['block_profiles'], g['blocks_per_rank']))
        for k in ('power', 'fields', 'jones'):
            x = g[k]
            print('   %-6s back-to-back %.2f ms, pipelined step %.2f ms (floor %.2f) check %s %s' % (k, x.get('ms', 0), x.get('pipelined_ms_per_step', 0), x.get('link_floor_ms', 0), x.get('pipelined_check'), x.get('error', '')))
    except Exception as e:
        print(tag, 'failed', e)
PY
head -20 gpurun_out/${T}_nccl_info_n$N.txt
