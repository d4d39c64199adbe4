#!/bin/bash
# Round-2 scaling check on an N-GPU box: bench.py at each N given on the command line (phase counters with PC_DEBUG=1 on rank 0)
set -u
mkdir -p gpurun_out
for N in "$@"; do
  PC_DEBUG=1 timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
      bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-configs --ensemble 0 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  tail -1 gpurun_out/scale_n$N.json | cut -c1-1500
  grep "pc dbg ms\|phase" gpurun_out/scale_n$N.err | tail -3
done
