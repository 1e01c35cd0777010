#!/bin/bash
# GPU call: bulk-path tests, then e2e with device scan vs host scan at all / 4 host threads
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest bulk subset"; timeout 900 python -m pytest tests -m gpu -x -q -k "decode_files or replicated or scan or damaged or chained or bulk" > $OUT/pytest_$TAG.log 2>&1; echo "exit $?"; tail -3 $OUT/pytest_$TAG.log
for HT in 0 4; do for GS in 1 0; do
echo "== e2e trace, host threads $HT, gpu_scan=$GS"; VPZ_TRACE=1 VPZ_BENCH_GPU_SCAN=$GS VPZ_BENCH_HOST_THREADS=$HT timeout 600 python bench.py --steps 5 --no-cpu --no-sub > $OUT/bench_${TAG}_t${HT}_scan$GS.json 2> $OUT/bench_${TAG}_t${HT}_scan$GS.err; echo "exit $?"; grep "vpz_decode_files" $OUT/bench_${TAG}_t${HT}_scan$GS.err | tail -1; python -c "import json; d=json.load(open('$OUT/bench_${TAG}_t${HT}_scan$GS.json')); e=d['e2e']; print('e2e %.2f G/s %.1f ms  s16 %.1f ms  link %.1f GB/s frac %.3f' % (e['value']/1e9, e['ms_per_step'], e['s16']['ms_per_step'], e.get('link_gbs_measured',0), e.get('frac_of_link',0)))"
done; done
