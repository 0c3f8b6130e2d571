"""Print the interesting keys of a bench.py JSON line (scratch helper for GPU sessions)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ("metric", "value", "ms_per_step", "value_l2_warm_back_to_back", "infer_events_per_sec", "critical_path_us", "mrr"):
    if k in d:
        print(k, "=", d[k])
if "e2e" in d:
    print("e2e =", d["e2e"]["value"])
for k in ("other_configs", "torch_gpu_baseline", "speedup_vs_torch_gpu", "drop_in_api", "drop_in_engine_loop", "cpu_baseline",
          "tgb_eval", "tgbl_wiki_999", "roofline"):
    if k in d:
        v = d[k]
        if isinstance(v, dict):
            v = {a: b for a, b in v.items() if a not in ("what", "sample", "api", "note", "how", "precision", "peak_source")}
        print(k, "=", v)
