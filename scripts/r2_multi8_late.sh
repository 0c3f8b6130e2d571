O=gpurun_out; mkdir -p $O
TRN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29611"
CTAN_EVAL_PREFETCH=2 timeout 300 $TRN --nproc-per-node 8 scripts/check_shard_equiv.py > $O/r2_equiv_late_n8.log 2>&1
echo "== equiv N=8 prefetch=2 rc=$?"; grep -E "shard-equivalence|Error|error" $O/r2_equiv_late_n8.log | head -5
CTAN_EVAL_PREFETCH=2 timeout 300 $TRN --nproc-per-node 8 bench.py --gpus 8 --steps 200 --warmup 10 --no-baselines > $O/late_pf2_n8.json 2> $O/late_pf2_n8.err
echo "== N=8 prefetch=2 rc=$?"; python scripts/show_bench.py $O/late_pf2_n8.json 2>/dev/null | head -4 | tr '\n' ' '; echo
