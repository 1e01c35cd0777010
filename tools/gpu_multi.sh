#!/bin/bash
# Multi-GPU call (gpurun --gpus N): BASELINE config 5 weak and strong, then the default bench line, under torchrun.
# Usage: gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi.sh N tag'
N=${1:-2}
TAG=${2:-r02}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nproc; nvidia-smi -L | head -8
for SC in weak strong; do
  echo "== config5 $SC, $N GPUs"
  VPZ_TRACE=1 timeout 600 $TR bench.py --gpus $N --workload config5 --scaling $SC --steps 10 --warmup 3 > $OUT/bench_${TAG}_config5_${SC}_${N}gpu.json 2> $OUT/bench_${TAG}_config5_${SC}_${N}gpu.err
  echo "exit $?"; python -c "import json; d=json.load(open('$OUT/bench_${TAG}_config5_${SC}_${N}gpu.json')); print('value %.2f G/s  %.1f ms/step  %.0f k excerpts/s' % (d['value']/1e9, d['ms_per_step'], d['config']['excerpts_per_s']/1e3))"
done
echo "== default bench, $N GPUs"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-sub > $OUT/bench_${TAG}_${N}gpu.json 2> $OUT/bench_${TAG}_${N}gpu.err
echo "exit $?"; python -c "import json; d=json.load(open('$OUT/bench_${TAG}_${N}gpu.json')); e=d['e2e']; print('value %.1f G/s %.2f ms | e2e %.2f G/s %.1f ms link %.1f GB/s frac %.3f | s16 %.1f ms' % (d['value']/1e9, d['ms_per_step'], e['value']/1e9, e['ms_per_step'], e.get('link_gbs_measured',0), e.get('frac_of_link',0), e['s16']['ms_per_step']))"
