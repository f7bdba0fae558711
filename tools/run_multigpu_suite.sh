#!/bin/bash
# Multi-GPU measurements of one round, run on an N-GPU box (N = first argument, default 8):
#   real-NCCL bit-parity (equal / uneven shards, eager / captured graph), the weak-scaling bench
#   with per-rank timing (equal and balanced shards, eager and graph), BASELINE configs[3] / [4]
#   as named, and the tracker-on-a-non-current-device test.  Everything lands in gpurun_out/.
N=${1:-8}
mkdir -p gpurun_out
TR="timeout 170 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29510
run() {  # name, nproc, command...
  local name=$1 np=$2; shift 2
  port=$((port + 1))
  echo "=== $name (nproc $np)"
  $TR --nproc-per-node $np --master-port $port "$@" > gpurun_out/mg_${name}.json 2> gpurun_out/mg_${name}.err
  echo "rc=$? $(tail -c 400 gpurun_out/mg_${name}.json)"
}
if [ "$N" -lt 8 ]; then run parity_${N}gpu $N tools/check_multigpu.py; fi
VPF_CHECK_SHARDS=uneven VPF_CHECK_GRAPH=1 run parity_${N}gpu_uneven_graph $N tools/check_multigpu.py
run bench_${N}gpu_equal $N bench.py --gpus $N --steps 8 --warmup 3
run bench_${N}gpu_balanced $N bench.py --gpus $N --steps 8 --warmup 3 --shards balanced
if [ "$N" -lt 8 ]; then run bench_${N}gpu_graph $N bench.py --gpus $N --steps 8 --warmup 3 --graph; fi
if [ "$N" -ge 8 ]; then
  for n in 2 4 8; do run cfg3_${n}gpu $n tools/bench_configs_multigpu.py --config cfg3 --steps 3 --warmup 2; done
  run cfg4_8gpu 8 tools/bench_configs_multigpu.py --config cfg4 --steps 3 --warmup 2
fi
timeout 120 python -m pytest tests/test_gpu_multirank.py -q -k non_current -p no:cacheprovider 2>&1 | tail -3
nvidia-smi --query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv > gpurun_out/mg_smi.csv
