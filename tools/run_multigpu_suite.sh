#!/bin/bash
# Multi-GPU measurements of one round on ONE 8-GPU box, packed into waves so the box time (charged
# 8x) stays short: runs that need fewer than 8 GPUs share a wave on disjoint GPU sets
# (CUDA_VISIBLE_DEVICES; power and clocks are per GPU).  Everything lands in gpurun_out/mg_*.
#   real-NCCL bit-parity at 8 GPUs (uneven shards, captured graph), the weak-scaling bench with
#   per-rank timing at N = 8 (equal and balanced shards) / 4 / 2 / 1, BASELINE configs[3] at 8 / 4 / 2 GPUs and configs[4] at 8.
mkdir -p gpurun_out
port=29510
launch() {  # name, gpu list, nproc, command...   (background; wait with `wait`)
  local name=$1 gpus=$2 np=$3; shift 3
  port=$((port + 1))
  ( CUDA_VISIBLE_DEVICES=$gpus timeout 120 python -m torch.distributed.run --nnodes=1 \
      --master-addr 127.0.0.1 --master-port $port --nproc-per-node $np "$@" \
      > gpurun_out/mg_${name}.json 2> gpurun_out/mg_${name}.err
    echo "=== $name (GPUs $gpus) rc=$? $(tail -c 300 gpurun_out/mg_${name}.json)" ) &
}
ALL=0,1,2,3,4,5,6,7
date +%T
launch bench_8gpu_equal $ALL 8 bench.py --gpus 8 --steps 8 --warmup 3; wait; date +%T
launch bench_8gpu_balanced $ALL 8 bench.py --gpus 8 --steps 8 --warmup 3 --shards balanced; wait; date +%T
VPF_CHECK_SHARDS=uneven VPF_CHECK_GRAPH=1 launch parity_8gpu_uneven_graph $ALL 8 tools/check_multigpu.py; wait; date +%T
launch cfg3_8gpu $ALL 8 tools/bench_configs_multigpu.py --config cfg3 --steps 3 --warmup 2; wait; date +%T
launch cfg4_8gpu $ALL 8 tools/bench_configs_multigpu.py --config cfg4 --steps 3 --warmup 2; wait; date +%T
launch bench_4gpu_equal 0,1,2,3 4 bench.py --gpus 4 --steps 8 --warmup 3
launch cfg3_4gpu 4,5,6,7 4 tools/bench_configs_multigpu.py --config cfg3 --steps 3 --warmup 2; wait; date +%T
launch bench_2gpu_equal 0,1 2 bench.py --gpus 2 --steps 8 --warmup 3
launch cfg3_2gpu 2,3 2 tools/bench_configs_multigpu.py --config cfg3 --steps 3 --warmup 2
# the same single-GPU bench on four different GPUs of the box: the GPU-to-GPU spread that the
# per-frame all-gather turns into waiting at N > 1
for g in 4 5 6 7; do launch bench_1gpu_dev$g $g 1 bench.py --gpus 1 --steps 8 --warmup 3 --no-extras; done
wait; date +%T
nvidia-smi --query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv > gpurun_out/mg_smi.csv
