#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), per-kernel timings, ncu launch list
# and one `--set full` capture of the dominant kernels.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh'
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit,memory.total --format=csv > $O/gpu.csv 2>&1
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
echo "== smoke"
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
tail -2 $O/smoke.log
echo "== bench"
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --no-fuse-uf --no-cpu-baseline > $O/bench_n1_unfused.json 2> $O/bench_n1_unfused.err; echo "bench unfused rc=$?"
cat $O/bench_n1.json $O/bench_n1_unfused.json $O/bench_ref.json
echo "== per-kernel"
timeout 600 python tools/bench_kernels.py > $O/kernels.log 2>&1; echo "kernels rc=$?"
timeout 600 python tools/bench_kernels.py --graph > $O/kernels_graph.log 2>&1; echo "kernels graph rc=$?"
grep -v '^{' $O/kernels_graph.log
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list (same command as the bench, 2 steps)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_bench.log 2>&1
echo "launches rc=$?"
echo "== ncu full"
N=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'head_kernel|head_uf|sweep|ufield|uf_' \
    -c 12 -f -o $O/prof_step python tools/run_once.py > $O/ncu_full.log 2>&1
echo "ncu full rc=$?"
fi
ls -la $O
