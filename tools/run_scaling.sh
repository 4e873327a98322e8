#!/bin/bash
# usage: tools/run_scaling.sh N [configs...]   one torchrun bench per config at N ranks, lines kept under gpurun_out/
N=$1; shift
mkdir -p gpurun_out
for c in "$@"; do
  SECONDS=0
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
    bench.py --gpus $N --config $c --steps 30 --warmup 5 > gpurun_out/r2_n${N}_$c.log 2> gpurun_out/r2_n${N}_$c.err
  echo "N=$N $c rc=$? ${SECONDS}s"
  tail -1 gpurun_out/r2_n${N}_$c.log | cut -c1-400
  tail -2 gpurun_out/r2_n${N}_$c.err | cut -c1-300
done
