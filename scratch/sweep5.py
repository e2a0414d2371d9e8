import sys
sys.path.insert(0,'/root/repo/scratch')
from acc_all import run
base = dict(n_inner_thin=5, ln_thin=0.2)
run('no10 ni7 hi12@.3', n_outer=10, n_inner=7, n_outer_hi=12, k_switch=0.3, **base)
run('no10 ni7 hi12@.45', n_outer=10, n_inner=7, n_outer_hi=12, k_switch=0.45, **base)
run('no10 ni6 hi12@.3', n_outer=10, n_inner=6, n_outer_hi=12, k_switch=0.3, **base)
run('no9 ni7 hi12@.3', n_outer=9, n_inner=7, n_outer_hi=12, k_switch=0.3, **base)
run('no10 ni7 hi12 dq.06', n_outer=10, n_inner=7, n_outer_hi=12, k_switch=0.3, bao_dq=0.06, **base)
