#!/bin/bash
# Extra GPU check: full-size ncu capture (traffic numbers).  compute-sanitizer is closed on this pool.
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
echo "== ncu full at 4096 streams"
CMD2="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
$CMD2 > $OUT/plain3_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:vpz_k -s 9 -c 3 -o $OUT/prof4096_$TAG -f $CMD2 > $OUT/ncu_full4096_$TAG.log 2>&1
echo "exit $?"; tail -2 $OUT/ncu_full4096_$TAG.log
