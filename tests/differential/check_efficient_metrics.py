"""Confusion-count metrics and Dice: ours vs the unmodified reference over a grid of batch / label axes, positive label,
threshold, background and discard settings, three streamed updates each (shapes and values of all four counts)."""
import itertools

import torch
from fl4health.metrics.efficient_metrics import BinaryDice as RB, MultiClassDice as RM
from fl4health.metrics.efficient_metrics_base import BinaryClassificationMetric as RBin, MultiClassificationMetric as RMul, ClassificationOutcome as RO
from fl4health_b200.metrics.efficient_metrics import BinaryDice as MB, MultiClassDice as MM
from fl4health_b200.metrics.efficient_metrics_base import BinaryClassificationMetric as MBin, MultiClassificationMetric as MMul, ClassificationOutcome as MO
torch.manual_seed(0)
n=0
def mk(cls):
    class C(cls):
        def compute_from_counts(self, true_positives, false_positives, true_negatives, false_negatives):
            return {"tp":true_positives,"fp":false_positives,"tn":true_negatives,"fn":false_negatives}
    return C
for batch_dim, label_dim, pos, thr, disc in itertools.product([None,0,1],[None,0,1,2],[0,1],[None,0.5],[None,"tn","tp","fp_fn"]):
    if batch_dim is not None and batch_dim==label_dim: continue
    shape=[4,3,5]
    if label_dim is not None: shape[label_dim]=2
    def d(O):
        return None if disc is None else {"tn":{O.TRUE_NEGATIVE},"tp":{O.TRUE_POSITIVE},"fp_fn":{O.FALSE_POSITIVE,O.FALSE_NEGATIVE}}[disc]
    r=mk(RBin)("m",label_dim=label_dim,batch_dim=batch_dim,pos_label=pos,threshold=thr,discard=d(RO))
    m=mk(MBin)("m",label_dim=label_dim,batch_dim=batch_dim,pos_label=pos,threshold=thr,discard=d(MO))
    for _ in range(3):
        p=torch.rand(shape); t=(torch.rand(shape)>0.5).float()
        r.update(p,t); m.update(p,t)
    a,b=r.compute(),m.compute()
    for k in a:
        assert a[k].shape==b[k].shape and torch.allclose(a[k].float(),b[k].float(),atol=1e-5),(batch_dim,label_dim,pos,thr,disc,k,a[k],b[k])
    n+=1
for batch_dim, label_dim, thr, bg, disc in itertools.product([None,0,2],[1],[None,0.5,1],[None,1],[None,"tn","fp_fn"]):
    def d(O):
        return None if disc is None else {"tn":{O.TRUE_NEGATIVE},"fp_fn":{O.FALSE_POSITIVE,O.FALSE_NEGATIVE}}[disc]
    r=mk(RMul)("m",label_dim=label_dim,batch_dim=batch_dim,threshold=thr,ignore_background=bg,discard=d(RO))
    m=mk(MMul)("m",label_dim=label_dim,batch_dim=batch_dim,threshold=thr,ignore_background=bg,discard=d(MO))
    for _ in range(3):
        p=torch.softmax(torch.randn(4,3,5),1); t=torch.nn.functional.one_hot(torch.randint(0,3,(4,5)),3).permute(0,2,1).float()
        r.update(p,t); m.update(p,t)
    a,b=r.compute(),m.compute()
    for k in a:
        assert a[k].shape==b[k].shape and torch.allclose(a[k].float(),b[k].float(),atol=1e-5),(batch_dim,thr,bg,disc,k)
    n+=1
for bd,pos,thr in itertools.product([None,0],[0,1],[None,0.5]):
    r=RB(batch_dim=bd,pos_label=pos,threshold=thr); m=MB(batch_dim=bd,pos_label=pos,threshold=thr)
    for _ in range(2):
        p=torch.rand(4,6,6); t=(torch.rand(4,6,6)>0.5).float(); r.update(p,t); m.update(p,t)
    a,b=r.compute("x"),m.compute("x"); assert a.keys()==b.keys() and all(abs(a[k]-b[k])<1e-6 for k in a),(a,b); n+=1
for bd,thr,bg in itertools.product([None,0],[None,1],[None,1]):
    r=RM(batch_dim=bd,label_dim=1,threshold=thr,ignore_background=bg); m=MM(batch_dim=bd,label_dim=1,threshold=thr,ignore_background=bg)
    for _ in range(2):
        p=torch.softmax(torch.randn(4,3,5),1); t=torch.nn.functional.one_hot(torch.randint(0,3,(4,5)),3).permute(0,2,1).float(); r.update(p,t); m.update(p,t)
    a,b=r.compute("x"),m.compute("x"); assert a.keys()==b.keys() and all(abs(a[k]-b[k])<1e-6 for k in a),(a,b); n+=1
print("configs agree:",n)
