OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-sub"
for K in k1a k1b; do
  ncu --set full --clock-control none --import-source on -k regex:vpz_$K -s 3 -c 1 -o $OUT/prof_p4_$K -f $CMD > $OUT/ncu_full_p4_$K.log 2>&1
  echo "$K exit $?"
done
