O=gpurun_out; mkdir -p $O
TRN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29611"
CTAN_EVAL_PREFETCH=2 timeout 600 $TRN --nproc-per-node 2 scripts/check_shard_equiv.py > $O/r2_equiv_late_n2.log 2>&1
echo "== equiv N=2 prefetch=2 rc=$?"; grep -E "shard-equivalence|Error|error" $O/r2_equiv_late_n2.log | head -5
for pf in 2 1; do
CTAN_EVAL_PREFETCH=$pf timeout 600 $TRN --nproc-per-node 2 bench.py --gpus 2 --steps 200 --warmup 10 --no-baselines > $O/late_pf${pf}_n2.json 2> $O/late_pf${pf}_n2.err
echo "== N=2 prefetch=$pf rc=$?"; python scripts/show_bench.py $O/late_pf${pf}_n2.json 2>/dev/null | head -4 | tr '\n' ' '; echo
done
