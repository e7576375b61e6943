import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import dropin_cases as dc
DEV='cuda:0'
keys=("total_loss","loss_box_cls","loss_box_reg","loss_mask")
rows=[]
for tag,kw in [("untouched",{}),("untouched again",{}),("1-ulp seed 1",dict(perturb=True,perturb_seed=1)),("1-ulp seed 2",dict(perturb=True,perturb_seed=2)),("1-ulp seed 3",dict(perturb=True,perturb_seed=3)),("1-ulp seed 4",dict(perturb=True,perturb_seed=4)),("1-ulp seed 5",dict(perturb=True,perturb_seed=5))]:
    m,_,i=dc.run_train_epoch(DEV, False, n_batches=2, **kw); print(tag, {k: round(m[k],5) for k in keys}, i["topk_ties_per_step"], flush=True)
m,_,i=dc.run_train_epoch(DEV, True, n_batches=2); print("patched", {k: round(m[k],5) for k in keys}, i["topk_ties_per_step"])
m,_,i=dc.run_train_epoch(DEV, True, n_batches=2); print("patched again", {k: round(m[k],5) for k in keys}, i["topk_ties_per_step"])
