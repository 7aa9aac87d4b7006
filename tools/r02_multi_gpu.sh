#!/bin/bash
# Multi-GPU evidence for round 2: run under `gpurun --gpus N -- 'bash tools/r02_multi_gpu.sh N [check]'`.
N=$1; O=gpurun_out/r02; mkdir -p $O
PORT=$((29500 + N))
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err
echo "bench N=$N rc=$?"; cut -c1-600 $O/bench_n$N.json
if [ "${2:-}" = "check" ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT + 1)) tools/multi_gpu_check.py > $O/multi_gpu_check_n$N.log 2> $O/multi_gpu_check_n$N.err
  echo "check N=$N rc=$?"; cat $O/multi_gpu_check_n$N.log | cut -c1-260
fi
if [ "${3:-}" = "ref" ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT + 2)) bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/bench_reference_n$N.json 2> $O/bench_reference_n$N.err
  echo "ref N=$N rc=$?"; cut -c1-400 $O/bench_reference_n$N.json
fi
