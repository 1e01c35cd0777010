#!/bin/bash
# Quick device-resident value check: stage parity subset + bench without e2e / cpu / sub-configs
TAG=${1:-v}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest stage subset"; timeout 600 python -m pytest tests -m gpu -x -q -k "stage_parity_every or trunc or batch_pcm or synth or reader_window" > $OUT/pytest_$TAG.log 2>&1; echo "exit $?"; tail -2 $OUT/pytest_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-sub > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "exit $?"
python -c "import json; d=json.load(open('$OUT/bench_$TAG.json')); print('value %.2f G/s ms/step %.2f | K1a %.2f K1b %.2f K3 %.2f' % (d['value']/1e9, d['ms_per_step'], d['roofline_k1a']['ms_per_launch'], d['roofline_k1b']['ms_per_launch'], d['roofline_k3']['ms_per_launch']))"
