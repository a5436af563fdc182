#!/bin/bash
# A/B of an experimental build against the committed library on the GPU box (ONE GPU):
#   here :  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC \
#                [-DEXPERIMENT_FLAG ...] -o variants/libX.so tamm_b200/csrc/tamm_kernels.cu
#   box  :  gpurun --timeout 600 -- 'tools/ab_variant.sh variants/libX.so X'
# Prints the short bench line (128 profiles, no e2e, no CPU arm) of both libraries and (unless AB_TESTS=0) runs the GPU
# parity suite on the variant.  The in-tree library is never touched: the variant is selected through TAMM_B200_LIB
# (tamm_b200/_lib.py), and the bench line records path + md5 of the library it loaded.  variants/ is scratch (*.so).
V=$(readlink -f "${1:?variant .so}"); T=${2:?tag}
mkdir -p gpurun_out
CMD="python bench.py --profiles 128 --steps 3 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/ab_base_$T.json 2> gpurun_out/ab_base_$T.err
TAMM_B200_LIB="$V" $CMD > gpurun_out/ab_var_$T.json 2> gpurun_out/ab_var_$T.err
if [ "${AB_TESTS:-1}" != "0" ]; then
  TAMM_B200_LIB="$V" python -m pytest tests -m gpu -x -q > gpurun_out/ab_tests_$T.log 2>&1; echo "tests rc=$?" >> gpurun_out/ab_tests_$T.log
fi
python - "$T" <<'PY'
import json, sys
t = sys.argv[1]
for k in ('base', 'var'):
    try:
        d = json.loads(open('gpurun_out/ab_%s_%s.json' % (k, t)).read().strip().splitlines()[-1])
        print(k, 'value %.4g  stack %.3f ms (%.1f %% of FP64 peak)  returns %.3f ms  lib %s' % (
            d['value'], d['roofline']['avg_launch_ms'], 100 * d['roofline']['frac'],
            (d.get('roofline_returns') or {}).get('avg_launch_ms', float('nan')), d.get('library', {}).get('md5')))
    except Exception as e:
        print(k, 'failed:', e)
PY
[ -f gpurun_out/ab_tests_$T.log ] && tail -2 gpurun_out/ab_tests_$T.log
