import sys
sys.path.insert(0,'/root/repo/scratch')
src = open('/root/repo/scratch/acc_default.py').read()
# round the tables to float32 before use
src = src.replace("kt = t['k']; Pk = t['plin_k']", "tabL = tabL.astype(np.float32).astype(np.float64)\ntabs = {k_: v.astype(np.float32).astype(np.float64) for k_, v in tabs.items()}\nkt = t['k']; Pk = t['plin_k']")
exec(src)
