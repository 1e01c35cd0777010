#!/bin/bash
# GPU call: the driver's two bench arms exactly as it runs them, plus timing of the whole default run
TAG=${1:-b}
OUT=gpurun_out
mkdir -p $OUT
echo "== bench (default)"; T0=$SECONDS; timeout 1200 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "exit $? after $((SECONDS-T0)) s"; tail -3 $OUT/bench_$TAG.err
python - <<PY
import json
d = json.load(open("$OUT/bench_$TAG.json"))
print("value %.2f G/s ms/step %.2f | K1a %.2f K1b %.2f K3 %.2f | e2e %.2f G/s %.1f ms link %.1f GB/s frac %.3f" % (d["value"]/1e9, d["ms_per_step"], d["roofline_k1a"]["ms_per_launch"], d["roofline_k1b"]["ms_per_launch"], d["roofline_k3"]["ms_per_launch"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["e2e"].get("link_gbs_measured",0), d["e2e"].get("frac_of_link",0)))
print("config3", d.get("config3",{}).get("roofline",{}).get("frac"), d.get("config3",{}).get("ms_per_step"))
print("config5", d.get("config5",{}).get("value"), d.get("config5",{}).get("ms_per_step"), d.get("config5",{}).get("cpu_baseline"))
print("cpu", d.get("cpu_baseline"))
PY
echo "== bench reference"; T0=$SECONDS; timeout 1200 python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "exit $? after $((SECONDS-T0)) s"; cat $OUT/bench_ref_$TAG.json | cut -c1-600
