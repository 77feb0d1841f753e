#!/bin/bash
# One GPU round (dev tool): full -m gpu suite, smoke, the contract bench (both arms), the other workloads,
# per-kernel graph timings, ncu launch list + one --set full capture of the step, compute-sanitizer.
# Outputs -> gpurun_out/   (copy what is to be kept into profiles/)
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.max.sm,power.limit --format=csv > $O/gpu.csv
python -m pytest tests -q -m gpu 2>&1 | tail -15 > $O/pytest_gpu.log; tail -2 $O/pytest_gpu.log
python -m pytest tests -q -m gpu 2>&1 | tail -3 > $O/pytest_gpu_again.log; tail -1 $O/pytest_gpu_again.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
for w in feedback upsample stress large_d stereo_refine feedback_refine; do
  python bench.py --workload $w --steps 50 > $O/bench_$w.json 2> $O/bench_$w.err; echo "$w rc=$?"; tail -2 $O/bench_$w.err
done
python tools/bench_kernels.py --graph > $O/kernels_graph.log 2>&1; grep -v "^{" $O/kernels_graph.log
for b in 2 8; do BASE3D_B=$b python tools/bench_kernels.py --only base3d 2>&1 | grep -v "^{" | tail -3; done > $O/base3d.log; cat $O/base3d.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_bench.log 2>&1; echo "ncu launch list rc=$?"
N=3 ncu --set full --clock-control none --import-source on -k regex:"sweep_xcorr|sweep_smaps|head_uf_tile" -s 5 -c 4 \
    -o $O/prof_step python tools/run_once.py > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"conv3x3" -s 4 -c 4 \
    -o $O/prof_conv python tools/bench_kernels.py --only refine > $O/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"conv3d_c32|conv3d_bn" -s 11 -c 4 \
    -o $O/prof_conv3d python tools/run_base3d.py > $O/ncu_conv3d.log 2>&1; echo "ncu conv3d rc=$?"
compute-sanitizer --tool memcheck python tools/run_small_step.py > $O/sanitizer_memcheck.log 2>&1; tail -3 $O/sanitizer_memcheck.log
compute-sanitizer --tool racecheck python tools/run_small_step.py > $O/sanitizer_racecheck.log 2>&1; tail -3 $O/sanitizer_racecheck.log
python - <<'PY'
import json
for n in ("n1", "feedback", "upsample", "stress", "large_d", "stereo_refine", "feedback_refine"):
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % n).read())
        print(n, "value %.0f  ms %.4f  roofline %.3f (in step %.3f)  frame_hbm %.3f  survey_hbm %.3f  e2e %.0f (%.2f of ceiling)  incumbent %s  cpu %s" % (
            d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("frac_in_step", 0), d["config"]["frame_hbm_frac"], d["config"]["survey_hbm_frac"],
            d["e2e"]["value"], d["e2e"]["frac_of_ceiling"], d.get("incumbent", {}).get("value"), d.get("cpu_baseline", {}).get("value")))
        print("   kernels", {k: round(v["ms"], 4) for k, v in d["kernels"].items()})
    except Exception as e:
        print(n, "failed", e)
PY
