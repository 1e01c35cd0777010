#!/bin/bash
# GPU call: full tests, bench default (trace), bench with host scan at 4 host threads vs device scan, config3
TAG=${1:-r02e}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "exit $?"; tail -5 $OUT/pytest_gpu_$TAG.log
echo "== bench (default)"; T0=$SECONDS; VPZ_TRACE=1 timeout 1200 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "exit $? after $((SECONDS-T0)) s"; grep "vpz_decode_files" $OUT/bench_$TAG.err | tail -2
python - <<PY
import json
d = json.load(open("$OUT/bench_$TAG.json"))
print("value %.2f G/s ms/step %.2f | K1a %.2f K1b %.2f K3 %.2f | e2e %.2f G/s %.1f ms link %.1f GB/s frac %.3f" % (d["value"]/1e9, d["ms_per_step"], d["roofline_k1a"]["ms_per_launch"], d["roofline_k1b"]["ms_per_launch"], d["roofline_k3"]["ms_per_launch"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["e2e"].get("link_gbs_measured",0), d["e2e"].get("frac_of_link",0)))
print("config3", d.get("config3",{}).get("roofline",{}).get("frac"), d.get("config3",{}).get("ms_per_step"))
print("config5", d.get("config5",{}).get("value"), d.get("config5",{}).get("ms_per_step"))
PY
for GS in 1 0; do
echo "== e2e trace, 4 host threads, gpu_scan=$GS"; VPZ_TRACE=1 VPZ_BENCH_GPU_SCAN=$GS VPZ_BENCH_HOST_THREADS=4 timeout 600 python bench.py --steps 3 --no-cpu --no-sub > $OUT/bench_${TAG}_scan$GS.json 2> $OUT/bench_${TAG}_scan$GS.err; echo "exit $?"; grep "vpz_decode_files" $OUT/bench_${TAG}_scan$GS.err | tail -1; python -c "import json; d=json.load(open('$OUT/bench_${TAG}_scan$GS.json')); print('e2e %.2f G/s %.1f ms' % (d['e2e']['value']/1e9, d['e2e']['ms_per_step']))"
done
