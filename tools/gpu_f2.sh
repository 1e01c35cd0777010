#!/bin/bash
TAG=${1:-f2}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest container + excerpts"; timeout 900 python -m pytest tests -m gpu -x -q -k "granule or excerpt or scan or damaged or chained or bulk" > $OUT/pytest_$TAG.log 2>&1; echo "exit $?"; tail -3 $OUT/pytest_$TAG.log
echo "== config5"; VPZ_TRACE=1 timeout 600 python bench.py --workload config5 --steps 10 --warmup 3 --no-cpu > $OUT/bench_${TAG}_config5.json 2> $OUT/bench_${TAG}_config5.err; echo "exit $?"; tail -2 $OUT/bench_${TAG}_config5.err
python -c "import json; d=json.load(open('$OUT/bench_${TAG}_config5.json')); print('value %.2f G/s  %.1f ms/step  %.0f k excerpts/s' % (d['value']/1e9, d['ms_per_step'], d['config']['excerpts_per_s']/1e3))"
