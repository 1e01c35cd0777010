#!/bin/bash
# Round-2 GPU call A: full GPU test suite, smoke, bench (own arm), ncu full capture of the three kernels.
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi_$TAG.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "exit $?"; tail -8 $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke_$TAG.log 2>&1; echo "exit $?"; tail -3 $OUT/smoke_$TAG.log
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
tail -3 $OUT/bench_$TAG.err
ls -la $OUT | tail -8
