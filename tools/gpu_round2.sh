#!/bin/bash
# N-GPU bench under torchrun (the driver's launch line) + reference arm at N.  Usage: gpu_round2.sh N
set -u
N=${1:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $O/gpus_n$N.csv
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 200 --warmup 10 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench N=$N rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 5 --warmup 1 > $O/bench_ref_n$N.json 2> $O/bench_ref_n$N.err; echo "ref N=$N rc=$?"
cat $O/bench_n$N.json $O/bench_ref_n$N.json
tail -5 $O/bench_n$N.err
