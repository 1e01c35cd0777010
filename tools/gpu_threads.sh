#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
nproc
for HT in 2 4 6 8 12 0; do
VPZ_TRACE=1 VPZ_BENCH_HOST_THREADS=$HT timeout 600 python bench.py --steps 6 --no-cpu --no-sub > $OUT/bench_ht$HT.json 2> $OUT/bench_ht$HT.err
python -c "import json; d=json.load(open('$OUT/bench_ht$HT.json')); e=d['e2e']; print('threads $HT: e2e %.2f G/s %.1f ms  s16 %.1f ms  link %.1f GB/s frac %.3f' % (e['value']/1e9, e['ms_per_step'], e['s16']['ms_per_step'], e.get('link_gbs_measured',0), e.get('frac_of_link',0)))"
grep "vpz_decode_files" $OUT/bench_ht$HT.err | tail -1
done
