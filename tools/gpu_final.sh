#!/bin/bash
# Round-end evidence on one GPU: full GPU suite, smoke, default bench line (everything in it), reference arm,
# config 3 line + its K3 DRAM traffic, ncu launch list, one full capture per kernel.
TAG=${1:-r02final}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "exit $?"; tail -3 $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke_$TAG.log 2>&1; echo "exit $?"; tail -1 $OUT/smoke_$TAG.log
echo "== bench (default)"; T0=$SECONDS; VPZ_TRACE=1 timeout 1500 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "exit $? after $((SECONDS-T0)) s"
python - <<PY
import json
d = json.load(open("$OUT/bench_$TAG.json"))
e = d["e2e"]
print("value %.2f G/s ms/step %.2f | K1a %.2f K1b %.2f K3 %.2f | e2e %.2f G/s %.1f ms link %.1f GB/s frac %.3f s16 %.1f ms | cpu %.2f G/s" % (d["value"]/1e9, d["ms_per_step"], d["roofline_k1a"]["ms_per_launch"], d["roofline_k1b"]["ms_per_launch"], d["roofline_k3"]["ms_per_launch"], e["value"]/1e9, e["ms_per_step"], e.get("link_gbs_measured",0), e.get("frac_of_link",0), e["s16"]["ms_per_step"], d["cpu_baseline"]["value"]/1e9))
print("config3 frac %.3f ms %.3f | config5 %.2f G/s %.1f ms cpu %.2f G/s" % (d["config3"]["roofline"]["frac"], d["config3"]["ms_per_step"], d["config5"]["value"]/1e9, d["config5"]["ms_per_step"], d["config5"]["cpu_baseline"]["value"]/1e9))
PY
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err; echo "exit $?"; python -c "import json; d=json.load(open('$OUT/bench_${TAG}_reference.json')); print('reference %.2f G/s on %d cores' % (d['value']/1e9, d['cpu_baseline']['cores']))"
echo "== config3 / config5 lines"; timeout 600 python bench.py --workload config3 > $OUT/bench_${TAG}_config3.json 2>/dev/null; timeout 600 python bench.py --workload config5 --steps 10 > $OUT/bench_${TAG}_config5.json 2>/dev/null; echo "exit $?"
echo "== ncu: config3 K3 traffic"
CMD3="python bench.py --workload config3 --steps 3 --warmup 3"
ncu --set full --clock-control none --import-source on -k regex:vpz_k3 -s 4 -c 1 -o $OUT/prof_${TAG}_config3 -f $CMD3 > $OUT/ncu_${TAG}_config3.log 2>&1; echo "exit $?"
echo "== ncu: launch list + one full capture per kernel"
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-sub"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-sub > $OUT/ncu_list_$TAG.log 2>&1; echo "list exit $?"
for K in k1a k1b k3; do
  ncu --set full --clock-control none --import-source on -k regex:vpz_$K -s 3 -c 1 -o $OUT/prof_${TAG}_$K -f $CMD > $OUT/ncu_full_${TAG}_$K.log 2>&1
  echo "$K exit $?"
done
