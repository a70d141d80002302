#!/bin/bash
# usage: tools/run_multi.sh NGPUS [extra bench args]   (GPU box; writes gpurun_out/bench_n$N.json)
N=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -15 gpurun_out/bench_n$N.err; cut -c1-1500 gpurun_out/bench_n$N.json
