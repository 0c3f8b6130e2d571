O=gpurun_out
TRN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29611"
for n in 4 8; do
for pf in 0 1; do
  CTAN_EVAL_PREFETCH=$pf timeout 600 $TRN --nproc-per-node $n bench.py --gpus $n --steps 200 --warmup 10 --no-baselines > $O/ab_pf${pf}_n$n.json 2> $O/ab_pf${pf}_n$n.err
  echo "== N=$n prefetch=$pf rc=$?"; python scripts/show_bench.py $O/ab_pf${pf}_n$n.json 2>&1 | head -5 | tr '\n' ' '; echo
done
done
