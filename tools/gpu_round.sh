#!/bin/bash
# One GPU round (dev tool): full -m gpu suite, smoke, the contract bench and the other workloads.  Outputs -> gpurun_out/
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.max.sm,power.limit --format=csv > $O/gpu.csv
python -m pytest tests -q -m gpu 2>&1 | tail -15 > $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
for w in feedback upsample stress large_d; do
  python bench.py --workload $w --steps 50 > $O/bench_$w.json 2> $O/bench_$w.err; echo "$w rc=$?"; tail -2 $O/bench_$w.err
done
python - <<'PY'
import json
for n in ("n1", "feedback", "upsample", "stress", "large_d"):
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % n).read())
        print(n, "value %.0f  ms %.4f  roofline %.3f  frame_hbm %.3f  survey_hbm %.3f  e2e %.0f (%.2f of ceiling)  incumbent %s  cpu %s" % (
            d["value"], d["ms_per_step"], d["roofline"]["frac"], d["config"]["frame_hbm_frac"], d["config"]["survey_hbm_frac"],
            d["e2e"]["value"], d["e2e"]["frac_of_ceiling"], d.get("incumbent", {}).get("value"), d.get("cpu_baseline", {}).get("value")))
        print("   kernels", {k: round(v["ms"], 4) for k, v in d["kernels"].items()})
    except Exception as e:
        print(n, "failed", e)
PY
