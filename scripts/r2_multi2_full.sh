O=gpurun_out; mkdir -p $O
TRN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29611"
for pf in default 0; do
if [ $pf = default ]; then unset CTAN_EVAL_PREFETCH; else export CTAN_EVAL_PREFETCH=$pf; fi
timeout 900 $TRN --nproc-per-node 2 bench.py --gpus 2 --steps 100 --warmup 10 > $O/full_n2_$pf.json 2> $O/full_n2_$pf.err
echo "== N=2 prefetch=$pf rc=$?"
python - <<PY
import json
d=json.loads(open("$O/full_n2_$pf.json").read().strip().splitlines()[-1])
w=d.get("tgbl_wiki_999") or {}
print({k:d.get(k) for k in ("n_gpus","value","ms_per_step","mrr")}, "wiki999", {k:w.get(k) for k in ("value","ms_per_step","mrr")})
PY
done
