import sys
sys.path.insert(0, '/root/repo/scratch'); sys.path.insert(0,'/root/repo')
from proto2 import *
for label, upd in [
    ('dq.06 n8', dict(bao_dq=0.06)),
    ('dq.06 n8 lnfar2 lnmid.9', dict(bao_dq=0.06, ln_far=2.0, ln_mid=0.9)),
    ('dq.045 n8 lnfar2 lnmid1', dict(bao_dq=0.045, ln_far=2.0, ln_mid=1.0)),
    ('dq.045 nin6', dict(bao_dq=0.045, n_inner=6)),
    ('dq.045 n7', dict(bao_dq=0.045, n_inner=7, n_outer=7)),
    ('dq.045 inner_scale1.5', dict(bao_dq=0.045, inner_scale=1.5)),
]:
    cfg = dict(base); cfg.update(upd)
    run(cfg, label)
