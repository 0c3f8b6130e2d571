#!/bin/bash
# ncu evidence of round 2 (1 GPU): launch list of the bench command, launch lists of eager training steps / evaluation
# batches, and a full capture of the step's dominant kernel (the tcgen05 contraction: qkvA forward and dX launches).
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-baselines"
$B > $O/r02_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench.csv $B > $O/r02_ncu_bench.log 2>&1
echo "== launches bench rc=$?"
S="python scripts/profile_step.py --steps 3"
$S > $O/r02_plain_step.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_step_wiki.csv $S > $O/r02_ncu_step.log 2>&1
echo "== launches step rc=$?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:qkva_tc -c 4 -o $O/r02_tc_wiki python scripts/profile_step.py --steps 2 > $O/r02_ncu_tc.log 2>&1
echo "== full capture tc rc=$?"
E="python scripts/profile_eval.py"
$E > $O/r02_plain_eval.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_eval_comment.csv $E > $O/r02_ncu_eval.log 2>&1
echo "== launches eval rc=$?"
ls -la $O | tail -20
