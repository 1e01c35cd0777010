#!/bin/bash
# GPU call: full tests, bench, ncu full capture of one launch of each kernel at the bench size, memcheck of the mixed batch
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "exit $?"; tail -5 $OUT/pytest_gpu_$TAG.log
echo "== bench"; VPZ_TRACE=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "exit $?"; python - <<PY
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
echo "== ncu full (4096 streams)"
CMD2="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
ncu --set full --clock-control none --import-source on -k regex:vpz_k -s 9 -c 3 -o $OUT/prof_$TAG -f $CMD2 > $OUT/ncu_full_$TAG.log 2>&1
echo "exit $?"; tail -2 $OUT/ncu_full_$TAG.log
echo "== memcheck mixed batch"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "mixed_batch and not fast" > $OUT/memcheck_$TAG.log 2>&1; echo "exit $?"; tail -5 $OUT/memcheck_$TAG.log
ls -la $OUT | tail -6
