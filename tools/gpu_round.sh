#!/bin/bash
# One gpurun call: GPU tests, smoke, bench (both arms), ncu launch list + one full capture.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi_$TAG.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "exit $?"; tail -5 $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke_$TAG.log 2>&1; echo "exit $?"; tail -3 $OUT/smoke_$TAG.log
echo "== bench"; VPZ_TRACE=1 timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "exit $?"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "exit $?"; cat $OUT/bench_ref_$TAG.json
echo "== ncu launch list"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --streams 1024"
$CMD > $OUT/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
echo "exit $?"; tail -2 $OUT/ncu_launch_$TAG.log
echo "== ncu full"
CMD2="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --streams 1024"
$CMD2 > $OUT/plain2_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:vpz_k -s 9 -c 3 -o $OUT/prof_$TAG -f $CMD2 > $OUT/ncu_full_$TAG.log 2>&1
echo "exit $?"; tail -2 $OUT/ncu_full_$TAG.log
ls -la $OUT
