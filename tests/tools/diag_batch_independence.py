import os, sys, subprocess, tempfile, numpy as np, pandas as pd
ROOT='/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'tests','tools'))
import make_random_ckpts, make_synth_dataset
d=tempfile.mkdtemp()
models=os.path.join(d,'ckpts'); make_random_ckpts.main(models, ["GCViTTiny-224x224"])
data=os.path.join(d,'data'); make_synth_dataset.main(data, 203)
res={}
for bs in (128, 102, 101, 100, 64, 203):
    out=os.path.join(d,f'o{bs}','pred.csv'); os.makedirs(os.path.dirname(out))
    env=dict(os.environ, VIP_MODEL_DIR=models, VIP_SAVE_PROBS='1', VIP_DEVICE_BATCH=str(bs))
    r=subprocess.run([sys.executable, os.path.join(ROOT,'main.py'), os.path.join(data,'input.csv'), out], env=env, capture_output=True, text=True)
    assert r.returncode==0, r.stderr[-2000:]
    res[bs]=pd.read_csv(os.path.join(d,f'o{bs}','temp','GCViTTiny-224x224_pred.csv')).logit.values
base=res[128]
for bs,v in res.items():
    diff=np.nonzero(v!=base)[0]
    print('batch', bs, 'differing images', len(diff), 'first', diff[:10], 'max abs', np.abs(v-base).max())
