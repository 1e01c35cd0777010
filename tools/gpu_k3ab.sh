#!/bin/bash
# K3 A/B on the GPU box: parity subset on the main build, then for libvpz.so and every vorbispizza_b200/alt_*.so
# (swapped in on the box only) the device-resident bench and the kernel-only config 3 bench.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_k3ab.sh tag'
TAG=${1:-k3ab}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest subset"; timeout 600 python -m pytest tests -m gpu -x -q -k "stage or trunc or batch_pcm or synth or decode_files or reader_window or s16 or general or mixed" > $OUT/pytest_$TAG.log 2>&1; echo "exit $?"; tail -3 $OUT/pytest_$TAG.log
one() {
  N=$1
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-sub > $OUT/bench_${TAG}_$N.json 2> $OUT/bench_${TAG}_$N.err; R1=$?
  timeout 300 python bench.py --workload config3 --no-cpu > $OUT/c3_${TAG}_$N.json 2> $OUT/c3_${TAG}_$N.err; R2=$?
  python - "$OUT/bench_${TAG}_$N.json" "$OUT/c3_${TAG}_$N.json" "$N" "$R1" "$R2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); c = json.load(open(sys.argv[2]))
    print("%-10s rc %s %s | step %.3f ms | K1a %.3f K1b %.3f K3 %.3f ms | config3 %.4f ms frac %.3f" % (
        sys.argv[3], sys.argv[4], sys.argv[5], d["ms_per_step"], d["roofline_k1a"]["ms_per_launch"], d["roofline_k1b"]["ms_per_launch"],
        d["roofline_k3"]["ms_per_launch"], c["ms_per_step"], c["roofline"]["frac"]))
except Exception as ex:
    print(sys.argv[3], "parse failed", ex)
PY
}
one main
cp vorbispizza_b200/libvpz.so /tmp/libvpz_main.so
for ALT in vorbispizza_b200/alt_*.so; do
  [ -f "$ALT" ] || continue
  cp $ALT vorbispizza_b200/libvpz.so
  one $(basename $ALT .so)
done
cp /tmp/libvpz_main.so vorbispizza_b200/libvpz.so
one main2
