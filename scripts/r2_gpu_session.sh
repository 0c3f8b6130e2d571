#!/bin/bash
# One GPU-box session of round 2: parity tests, then the bench (+ optional A/B env toggles given as arguments, e.g.
# `bash scripts/r2_gpu_session.sh CTAN_TC_ROW_HINT=0`).  Everything lands in gpurun_out/ (scratch).
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -X faulthandler -m pytest tests -m gpu -q > $O/r2_tests.log 2>&1
echo "== tests rc=$?"; grep -E "passed|failed" $O/r2_tests.log | tail -3; grep -E "^(FAILED|ERROR)" $O/r2_tests.log | head -20
show() {
python - <<PY
import json
d = json.loads(open("$1").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "value_l2_warm_back_to_back", "infer_events_per_sec", "critical_path_us", "e2e"):
    print(k, "=", d.get(k))
for k in ("other_configs", "torch_gpu_baseline", "speedup_vs_torch_gpu", "drop_in_api", "drop_in_engine_loop", "cpu_baseline", "tgb_eval"):
    if k in d: print(k, "=", d.get(k))
PY
}
timeout 300 python __graft_entry__.py smoke > $O/r2_smoke.log 2>&1
echo "== smoke rc=$?"; tail -2 $O/r2_smoke.log
for ab in "$@"; do
  env $ab timeout 300 python bench.py --steps 100 --warmup 10 --no-baselines > $O/r2_bench_ab.json 2> $O/r2_bench_ab.err
  echo "== A/B bench with $ab rc=$?"; show $O/r2_bench_ab.json
done
timeout 900 python bench.py --steps 100 --warmup 10 > $O/r2_bench.json 2> $O/r2_bench.err
echo "== bench rc=$?"; tail -c 600 $O/r2_bench.err; show $O/r2_bench.json
