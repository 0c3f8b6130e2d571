#!/bin/bash
# One GPU-box session of round 2: parity tests, then the bench.  Everything lands in gpurun_out/ (scratch); what is worth
# keeping is copied to profiles/ afterwards.
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -X faulthandler -m pytest tests -m gpu -q > $O/r2_tests.log 2>&1
echo "== tests rc=$?"; grep -E "passed|failed" $O/r2_tests.log | tail -3; grep -E "^(FAILED|ERROR)" $O/r2_tests.log | head -20
timeout 600 python bench.py --steps 100 --warmup 10 > $O/r2_bench.json 2> $O/r2_bench.err
echo "== bench rc=$?"; tail -c 600 $O/r2_bench.err
python - <<PY
import json
d = json.loads(open("$O/r2_bench.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "value_l2_warm_back_to_back", "infer_events_per_sec", "critical_path_us", "e2e", "other_configs",
          "torch_gpu_baseline", "speedup_vs_torch_gpu", "drop_in_api", "drop_in_engine_loop", "cpu_baseline", "tgb_eval"):
    print(k, "=", d.get(k))
PY
