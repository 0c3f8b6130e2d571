#!/bin/bash
# 8-GPU box: shard equivalence at 8 ranks with the front-end prefetch on, then the prefetch A/B at 8 and the series 2/4/8.
O=gpurun_out; mkdir -p $O
TRN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29611"
CTAN_EVAL_PREFETCH=1 timeout 600 $TRN --nproc-per-node 8 scripts/check_shard_equiv.py > $O/r2_equiv_peer_pf1_n8.log 2>&1
echo "== equiv N=8 prefetch=1 rc=$?"; grep -E "shard-equivalence|Error|error" $O/r2_equiv_peer_pf1_n8.log | head -5
run() {  # n pf
CTAN_EVAL_PREFETCH=$2 timeout 600 $TRN --nproc-per-node $1 bench.py --gpus $1 --steps 200 --warmup 10 --no-baselines > $O/pf$2_n$1.json 2> $O/pf$2_n$1.err
echo "== N=$1 prefetch=$2 rc=$?"; python scripts/show_bench.py $O/pf$2_n$1.json 2>/dev/null | head -4 | tr '\n' ' '; echo
}
run 8 0
run 8 1
run 4 1
run 4 0
run 2 1
