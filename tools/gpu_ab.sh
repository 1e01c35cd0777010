#!/bin/bash
# A/B iteration on the GPU box: parity subset, short bench of libvpz.so and of every alternative build
# vorbispizza_b200/alt_*.so (swapped in on the box only), then one full ncu capture of the main build.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_ab.sh tag [kernel-regex] [pytest -k expr]'
TAG=${1:-ab}
KRE=${2:-vpz_k}
KEXPR=${3:-"stage or batch_pcm or synth or decode_files or trunc"}
OUT=gpurun_out
SKIP=6; [ "$KRE" != "vpz_k" ] && SKIP=3
mkdir -p $OUT
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    e = d.get("e2e", {})
    print("value %.3f G/s  ms/step %.2f | K1a %.2f ms | K1b %.2f ms frac %.3f | K3 %.2f ms %.0f GB/s frac %.3f | e2e %.3f G/s %.1f ms" % (
        d["value"]/1e9, d["ms_per_step"], d["roofline_k1a"]["ms_per_launch"], d["roofline_k1b"]["ms_per_launch"], d["roofline_k1b"]["frac"],
        d["roofline_k3"]["ms_per_launch"], d["roofline_k3"]["achieved"], d["roofline_k3"]["frac"], e.get("value", 0)/1e9, e.get("ms_per_step", 0)))
except Exception as ex:
    print("bench parse failed", ex)
PY
}
echo "== pytest subset"; timeout 600 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $OUT/pytest_$TAG.log 2>&1; echo "exit $?"; tail -4 $OUT/pytest_$TAG.log
echo "== bench main"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "exit $?"; summ $OUT/bench_$TAG.json; tail -2 $OUT/bench_$TAG.err
cp vorbispizza_b200/libvpz.so /tmp/libvpz_main.so
for ALT in vorbispizza_b200/alt_*.so; do
  [ -f "$ALT" ] || continue
  N=$(basename $ALT .so)
  cp $ALT vorbispizza_b200/libvpz.so
  echo "== bench $N"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > $OUT/bench_${TAG}_$N.json 2> $OUT/bench_${TAG}_$N.err; echo "exit $?"; summ $OUT/bench_${TAG}_$N.json; tail -2 $OUT/bench_${TAG}_$N.err
done
cp /tmp/libvpz_main.so vorbispizza_b200/libvpz.so
# extra bench variants of the main build: BENCH_VARIANTS="--l1-bits 8;--l1-bits 10"
IFS=';' read -ra VARS <<< "$BENCH_VARIANTS"
for V in "${VARS[@]}"; do
  [ -n "$V" ] || continue
  N=$(echo $V | tr -c 'a-zA-Z0-9' '_')
  echo "== bench $V"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e $V > $OUT/bench_${TAG}_$N.json 2> $OUT/bench_${TAG}_$N.err; echo "exit $?"; summ $OUT/bench_${TAG}_$N.json; tail -2 $OUT/bench_${TAG}_$N.err
done
if [ "$KRE" != "none" ]; then
echo "== ncu full"
CMD2="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -o $OUT/prof_$TAG -f $CMD2 > $OUT/ncu_full_$TAG.log 2>&1
echo "exit $?"; tail -2 $OUT/ncu_full_$TAG.log
fi
