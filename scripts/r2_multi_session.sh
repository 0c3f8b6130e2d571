#!/bin/bash
# Multi-GPU session: bitwise shard equivalence at N ranks, then the scaling series 1..N of the sharded-evaluation bench
# (peer-memory exchange), plus the NCCL exchange at N for comparison.   usage: bash scripts/r2_multi_session.sh <N>
N=${1:-2}
O=gpurun_out
mkdir -p $O
show() {
python - <<PY
import json
try:
    d = json.loads(open("$1").read().strip().splitlines()[-1])
    w = d.get("tgbl_wiki_999") or {}
    print({k: d.get(k) for k in ("n_gpus", "value", "ms_per_step", "mrr", "launches_per_step")}, "e2e", d["e2e"]["value"],
          "| wiki999", {k: w.get(k) for k in ("value", "ms_per_step", "mrr", "rank0_subgraph")})
except Exception as e:
    print("parse failed", e); print(open("$1".replace(".json", ".err")).read()[-2500:])
PY
}
TRN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TRN --nproc-per-node $N scripts/check_shard_equiv.py > $O/r2_equiv_peer_n$N.log 2>&1
echo "== equiv exchange=peer N=$N rc=$?"; grep -E "shard-equivalence|Error|error" $O/r2_equiv_peer_n$N.log | head -5
timeout 600 python bench.py --gpus 1 --metric tgb_eval --steps 100 --warmup 10 > $O/r2_scale_peer_n1.json 2> $O/r2_scale_peer_n1.err
echo "== bench N=1 rc=$?"; show $O/r2_scale_peer_n1.json
for n in 2 4 8; do
  [ $n -le $N ] || continue
  timeout 900 $TRN --nproc-per-node $n bench.py --gpus $n --steps 100 --warmup 10 > $O/r2_scale_peer_n$n.json 2> $O/r2_scale_peer_n$n.err
  echo "== bench exchange=peer N=$n rc=$?"; show $O/r2_scale_peer_n$n.json
done
CTAN_EVAL_EXCHANGE=nccl timeout 900 $TRN --nproc-per-node $N bench.py --gpus $N --steps 100 --warmup 10 > $O/r2_scale_nccl_n$N.json 2> $O/r2_scale_nccl_n$N.err
echo "== bench exchange=nccl N=$N rc=$?"; show $O/r2_scale_nccl_n$N.json
