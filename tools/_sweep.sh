for rep in 1 2; do for v in A B C; do echo -n "$v: "; timeout 300 python tools/prof_aux.py --name quadrotor_$v --batch 4096 --reps 4 2>&1 | tail -1 | cut -c30-110; done; done
