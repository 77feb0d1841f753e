#!/bin/bash
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -m gpu -x -k "sweep or large_d or frame" 2>&1 | tail -4
python tools/bench_kernels.py --graph --only fuse 2>&1 | grep -v "^{"
python tools/bench_kernels.py --graph --only fuse --pose mono 2>&1 | grep -v "^{"
python bench.py --workload stress --steps 20 --no-cpu-baseline > $O/bench_stress.json 2> $O/bench_stress.err; tail -2 $O/bench_stress.err
python bench.py --steps 100 --no-cpu-baseline > $O/bench_quick.json 2> $O/bench_quick.err; tail -2 $O/bench_quick.err
python - <<'PY'
import json
for n in ("quick", "stress"):
    d = json.loads(open("gpurun_out/bench_%s.json" % n).read())
    print(n, "value %.0f  ms %.4f  roofline %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"]), {k: round(v["ms"], 4) for k, v in d["kernels"].items()}, d.get("sweep", {}).get("fp32_frac_direct_form"))
PY
