#!/bin/bash
# head + UF kernel variants (dev tool)
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -m gpu -k "uf or head or frame or pipeline" 2>&1 | tail -3
for cfg in "0" "1"; do
  echo "== prefetch=$cfg"
  DPV_UFTILE_PREFETCH=$cfg python tools/bench_kernels.py --graph --only ufield 2>&1 | grep -E "head_full_uf_fused|head_full_groundplane" | grep -v "^{"
done
python bench.py --no-cpu-baseline --steps 100 > $O/bench_quick.json; python -c "
import json; d=json.loads(open('$O/bench_quick.json').read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['kernels'])"
