#!/bin/bash
# Quick GPU iteration: parity subset, bench without the CPU leg, one full ncu capture.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_quick.sh tag [kernel-regex] [pytest -k expr]'
TAG=${1:-q}
KRE=${2:-vpz_k}
KEXPR=${3:-"batch_pcm or synth or decode_files"}
OUT=gpurun_out
SKIP=6; [ "$KRE" != "vpz_k" ] && SKIP=3
mkdir -p $OUT
echo "== pytest subset"; timeout 600 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $OUT/pytest_$TAG.log 2>&1; echo "exit $?"; tail -4 $OUT/pytest_$TAG.log
echo "== bench"; VPZ_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "exit $?"; python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$TAG.json"))
    print("value %.3f G/s  ms/step %.2f | K1a %.2f ms | K1b %.2f ms frac %.3f | K3 %.2f ms %.0f GB/s frac %.3f | e2e %.3f G/s %.1f ms" % (
        d["value"]/1e9, d["ms_per_step"], d["roofline_k1a"]["ms_per_launch"], d["roofline_k1b"]["ms_per_launch"], d["roofline_k1b"]["frac"],
        d["roofline_k3"]["ms_per_launch"], d["roofline_k3"]["achieved"], d["roofline_k3"]["frac"],
        d.get("e2e",{}).get("value",0)/1e9, d.get("e2e",{}).get("ms_per_step",0)))
except Exception as e:
    print("bench parse failed", e)
PY
tail -2 $OUT/bench_$TAG.err
echo "== ncu full"
CMD2="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --streams 1024"
$CMD2 > $OUT/plain2_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 2 -o $OUT/prof_$TAG -f $CMD2 > $OUT/ncu_full_$TAG.log 2>&1
echo "exit $?"; tail -2 $OUT/ncu_full_$TAG.log
