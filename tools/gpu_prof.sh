#!/bin/bash
# One full ncu capture per kernel at the bench size (4,096 streams), plus the launch list.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-sub"
$CMD > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
for K in k1a k1b k3; do
  ncu --set full --clock-control none --import-source on -k regex:vpz_$K -s 3 -c 1 -o $OUT/prof_${TAG}_$K -f $CMD > $OUT/ncu_full_${TAG}_$K.log 2>&1
  echo "$K exit $?"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_list_$TAG.log 2>&1
echo "list exit $?"
