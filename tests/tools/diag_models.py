import numpy as np, torch, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import resnet_rs as R, gcvit as G, preprocess as P
from vipcup_b200.models import ResNetRS, GCViT
dev=torch.device('cuda:0')
def report(name, ref_taps, taps, Wk, Wb):
    for k in ref_taps:
        a=taps[k].float().cpu().numpy(); b=ref_taps[k]
        print(f"  {k:8s} scale {np.abs(b).mean():.3f} rms err {np.sqrt(((a-b)**2).mean()):.4e} max {np.abs(a-b).max():.3e}  inter-image std {b.std(axis=0).mean():.4f}")
    fg=taps['feat'].cpu().numpy(); fr=ref_taps['feat']
    lg=fg@Wk+Wb; lr=fr@Wk+Wb
    print(f"  {name}: logits ref std {lr.std():.3f} range [{lr.min():.2f},{lr.max():.2f}] max abs logit err {np.abs(lg-lr).max():.4e}")
n=16
x200=np.stack([P.decode_to_float(P.synth_image(i),200,200) for i in range(n)])
x224=np.stack([P.decode_to_float(P.synth_image(i),224,224) for i in range(n)])
W=R.random_weights(50,2,seed=3); rt={}; R.forward(x200,W,50,taps=rt)
m=ResNetRS(50,classes=2,device=dev).load_weights(W); t={}; m(torch.from_numpy(x200).to(dev),taps=t); report("rs50",rt,t,W['predictions/kernel'],W['predictions/bias'])
for v in ["tiny","small"]:
    W=G.random_weights(v,2,seed=5); rt={}; G.forward(x224,W,v,taps=rt)
    m=GCViT(v,num_classes=2,device=dev).load_weights(W); t={}; m(torch.from_numpy(x224).to(dev),taps=t); report("gcvit-"+v,rt,t,W['head/kernel'],W['head/bias'])
