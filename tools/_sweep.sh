for n in quadrotor quadrotor; do echo -n "$n: "; timeout 300 python tools/prof_aux.py --name $n --batch 4096 --reps 4 2>&1 | tail -1 | cut -c30-110; done
for b in 148 1480; do echo -n "batch $b: "; timeout 300 python tools/prof_aux.py --batch $b --reps 4 2>&1 | tail -1 | cut -c30-110; done
