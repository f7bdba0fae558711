#!/bin/bash
# Weak scaling of the headline workload on ONE 8-GPU box, in waves that keep the 8x-charged box time
# short: N = 8; N = 4 beside four single-GPU runs (the box's GPU-to-GPU spread); N = 2; then BASELINE
# configs[3] and configs[4] on 8 GPUs.  Results: gpurun_out/mg_*.json
mkdir -p gpurun_out
port=29610
launch() {  # name, gpu list, nproc, command...
  local name=$1 gpus=$2 np=$3; shift 3
  port=$((port + 1))
  ( CUDA_VISIBLE_DEVICES=$gpus timeout 120 python -m torch.distributed.run --nnodes=1 \
      --master-addr 127.0.0.1 --master-port $port --nproc-per-node $np "$@" \
      > gpurun_out/mg_${name}.json 2> gpurun_out/mg_${name}.err
    echo "=== $name (GPUs $gpus) rc=$? $(tail -c 200 gpurun_out/mg_${name}.json)" ) &
}
ALL=0,1,2,3,4,5,6,7
date +%T
launch bench_8gpu_equal $ALL 8 bench.py --gpus 8 --steps 8 --warmup 3; wait; date +%T
launch bench_4gpu_equal 0,1,2,3 4 bench.py --gpus 4 --steps 8 --warmup 3
for g in 4 5 6 7; do launch bench_1gpu_dev$g $g 1 bench.py --gpus 1 --steps 8 --warmup 3 --no-extras; done
wait; date +%T
launch bench_2gpu_equal 0,1 2 bench.py --gpus 2 --steps 8 --warmup 3
for g in 2 3; do launch bench_1gpu_dev$g $g 1 bench.py --gpus 1 --steps 8 --warmup 3 --no-extras; done
wait; date +%T
launch cfg3_8gpu $ALL 8 tools/bench_configs_multigpu.py --config cfg3 --steps 3 --warmup 2; wait; date +%T
launch cfg4_8gpu $ALL 8 tools/bench_configs_multigpu.py --config cfg4 --steps 3 --warmup 2; wait; date +%T
