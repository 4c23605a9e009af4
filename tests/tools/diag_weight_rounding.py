import sys, numpy as np, torch
sys.path.insert(0,'' + __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))) + '')
from oracle import gcvit as G, preprocess as P
def bf(a): return torch.from_numpy(a).to(torch.bfloat16).float().numpy()
for variant, seed, hw in (("tiny",7,200),("small",3,224)):
    W=G.random_weights(variant,2,seed=seed)
    x=np.stack([P.decode_to_float(P.synth_image(i),hw,hw) for i in range(8)])
    taps={}; ref=G.forward(x,W,variant,head_act="softmax",taps=taps) if 'taps' in G.forward.__code__.co_varnames else None
    def logits(Wq):
        t={}; G.forward(x,Wq,variant,head_act="softmax",taps=t); return t["feat"]@W["head/kernel"]+W["head/bias"]
    l0=logits(W)
    groups={"all kernels":lambda k:k.endswith("kernel") and not k.startswith("head"),
            "block kernels (levels/*/blocks)":lambda k:k.endswith("kernel") and "/blocks/" in k,
            "non-block kernels (stem, downsample, q-gen)":lambda k:k.endswith("kernel") and "/blocks/" not in k and not k.startswith("head"),
            "main-path convs only":lambda k:k.endswith("kernel") and "/blocks/" not in k and not k.startswith("head") and "to_q_global" not in k and "q_global" not in k}
    print(variant, "logit std over images", l0.std(0))
    for name,sel in groups.items():
        Wq={k:(bf(v) if sel(k) else v) for k,v in W.items()}
        n=sum(sel(k) for k in W)
        print(f"  bf16-rounded {name:45s} ({n:3d} tensors): max |dlogit| {np.abs(logits(Wq)-l0).max():.3e}")
    xq=bf(x); t={}; G.forward(xq,W,variant,head_act="softmax",taps=t)
    print(f"  bf16-rounded input image: max |dlogit| {np.abs(t['feat']@W['head/kernel']+W['head/bias']-l0).max():.3e}")
