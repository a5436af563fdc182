#!/bin/bash
# One bench line per named configuration of BASELINE.json (configs[0..4]) on ONE GPU, plus the alpha = 0 variant of
# configs[1] (normal incidence: the closed-form eigen path).  Lines land in gpurun_out/bench_cfg<i>_<TAG>.json.
#   gpurun --timeout 900 -- 'tools/bench_all_configs.sh r02'
T=${1:?tag}
mkdir -p gpurun_out
for c in 0 1 2 3; do
  python bench.py --config $c --steps 20 --warmup 3 --cpu-seconds 5 > gpurun_out/bench_cfg${c}_$T.json 2> gpurun_out/bench_cfg${c}_$T.err || echo "config $c failed"
done
python bench.py --config 1 --alpha-deg 0 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_cfg1_a0_$T.json 2> gpurun_out/bench_cfg1_a0_$T.err || echo "config 1 alpha 0 failed"
python - "$T" <<'PY'
import json, sys
t = sys.argv[1]
for c in ('0', '1', '1_a0', '2', '3'):
    try:
        d = json.loads(open('gpurun_out/bench_cfg%s_%s.json' % (c, t)).read().strip().splitlines()[-1])
        e = d.get('e2e') or {}
        print('cfg%-5s value %.4g evals/s  %.3f ms/step  stack %.1f %% of FP64 peak  e2e %s evals/s (%.3f ms/call)  cpu %s' % (
            c, d['value'], d['ms_per_step'], 100 * d['roofline']['frac'], ('%.4g' % e['value']) if e.get('value') else e.get('error'),
            e.get('ms_per_call', float('nan')), ('%.4g (%s, %d cores)' % (d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores'])) if 'cpu_baseline' in d else '-'))
    except Exception as ex:
        print('cfg', c, 'failed:', ex)
PY
