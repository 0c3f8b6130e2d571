#!/bin/bash
# Multi-GPU session: bitwise shard equivalence, then the sharded-evaluation bench with the exchange variants.
# usage: bash scripts/r2_multi_session.sh <N>
N=${1:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
for ex in peer nccl; do
  CTAN_EVAL_EXCHANGE=$ex timeout 600 $TR scripts/check_shard_equiv.py > $O/r2_equiv_${ex}_n$N.log 2>&1
  echo "== equiv exchange=$ex N=$N rc=$?"; grep -E "shard-equivalence|Error|error" $O/r2_equiv_${ex}_n$N.log | head -5
done
for cfg in "peer 0" "peer 1" "nccl 0"; do
  set -- $cfg
  CTAN_EVAL_EXCHANGE=$1 CTAN_EVAL_PREFETCH=$2 timeout 900 $TR bench.py --gpus $N --steps 100 --warmup 10 > $O/r2_scale_$1_pf$2_n$N.json 2> $O/r2_scale_$1_pf$2_n$N.err
  echo "== bench exchange=$1 prefetch=$2 N=$N rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$O/r2_scale_$1_pf$2_n$N.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "mrr", "launches_per_step", "sharding")}, "e2e", d["e2e"]["value"], "wiki999", d.get("tgbl_wiki_999"))
except Exception as e:
    print("parse failed", e); print(open("$O/r2_scale_$1_pf$2_n$N.err").read()[-2500:])
PY
done
