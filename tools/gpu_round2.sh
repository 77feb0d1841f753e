#!/bin/bash
# N-GPU round under torchrun (the driver's launch line).  Usage: gpu_round2.sh N [workloads...]
set -u
N=${1:-2}; shift
WLS=${@:-"stereo large_d upsample feedback"}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $O/gpus_n$N.csv
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29510 tools/nccl_plane_shard.py > $O/nccl_plane_shard_n$N.txt 2> $O/nccl_plane_shard_n$N.err; echo "plane shard N=$N rc=$?"; cat $O/nccl_plane_shard_n$N.txt
for w in $WLS; do
  steps=200
  timeout 900 $TR --master-port 29511 bench.py --gpus $N --workload $w --steps $steps --warmup 10 > $O/bench_${w}_n$N.json 2> $O/bench_${w}_n$N.err; echo "bench $w N=$N rc=$?"
  tail -2 $O/bench_${w}_n$N.err | cut -c1-300
done
timeout 600 $TR --master-port 29512 bench.py --impl reference --gpus $N --steps 5 --warmup 1 > $O/bench_ref_n$N.json 2> $O/bench_ref_n$N.err; echo "ref N=$N rc=$?"
python - <<PY
import json
for w in "$WLS".split():
    try:
        d = json.loads(open("gpurun_out/bench_%s_n$N.json" % w).read())
        print(w, "N=%d value %.0f  ms %.4f  roofline %.3f  e2e %.0f (%.2f of ceiling %.0f)" % (d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["frac_of_ceiling"], d["e2e"]["ceiling"]["value"]))
    except Exception as e:
        print(w, "failed", e)
PY
