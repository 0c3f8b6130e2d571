#!/bin/bash
# ncu evidence of the final round-2 step (1 GPU): launch lists (bench command, eager training steps, evaluation batches),
# DRAM bytes of every kernel of a training step, and a full capture of the kernels on the step's critical path.
# Every ncu command runs only after the same command exited 0 without ncu.
O=gpurun_out
mkdir -p $O
S="python scripts/profile_step.py --steps 3"
$S > $O/r02_plain_step.log 2>&1 || { echo "plain step failed"; tail -5 $O/r02_plain_step.log; exit 1; }
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_step_wiki_v3.csv $S > $O/r02_ncu_step.log 2>&1
echo "== launches step rc=$?"
ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_dram_step_wiki.csv python scripts/profile_step.py --steps 2 > $O/r02_ncu_dram.log 2>&1
echo "== dram step rc=$?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:edge_bwd|edge_fwd|link_fused|qkva_tc|fuse_enc_split|eproj_rows" -c 7 -o $O/r02_step_full python scripts/profile_step.py --steps 1 > $O/r02_ncu_full.log 2>&1
echo "== full capture rc=$?"
ncu -i $O/r02_step_full.ncu-rep --page raw --csv > $O/r02_step_full_raw.csv 2>/dev/null
ncu -i $O/r02_step_full.ncu-rep --page details > $O/r02_step_full_details.txt 2>/dev/null
B="python bench.py --steps 2 --warmup 3 --no-baselines"
$B > $O/r02_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench_v3.csv $B > $O/r02_ncu_bench.log 2>&1
echo "== launches bench rc=$?"
E="python scripts/profile_eval.py"
$E > $O/r02_plain_eval.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_eval_comment_v3.csv $E > $O/r02_ncu_eval.log 2>&1
echo "== launches eval rc=$?"
ls -la $O | tail -16
