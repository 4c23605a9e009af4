"""Which images differ between a 1-GPU and a 2-GPU run of main.py, and by how much (diagnostic for the bit-reproducibility
test)."""
import os, sys, subprocess, tempfile, numpy as np, pandas as pd
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'tools'))
import make_random_ckpts, make_synth_dataset
d = tempfile.mkdtemp()
names = ["GCViTTiny-224x224", "ResNetRS50-200x200"]
models = os.path.join(d, 'ckpts'); make_random_ckpts.main(models, names)
data = os.path.join(d, 'data'); make_synth_dataset.main(data, 203)
res = {}
for tag, nproc, extra in (("1gpu", 1, {}), ("2gpu", 2, {}), ("1gpu_omp1", 1, {"OMP_NUM_THREADS": "1"})):
    out = os.path.join(d, tag, 'pred.csv'); os.makedirs(os.path.dirname(out))
    env = dict(os.environ, VIP_MODEL_DIR=models, VIP_SAVE_PROBS='1', **extra)
    cmd = [sys.executable, os.path.join(ROOT, 'main.py')]
    if nproc > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
               "--master-port", "29541", os.path.join(ROOT, 'main.py')]
    r = subprocess.run(cmd + [os.path.join(data, 'input.csv'), out], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    res[tag] = {m: pd.read_csv(os.path.join(d, tag, 'temp', m + '_pred.csv')) for m in names}
for m in names:
    base = res["1gpu"][m]
    for tag in res:
        v = res[tag][m]
        assert list(v.filename) == list(base.filename)
        diff = np.nonzero(v.logit.values != base.logit.values)[0]
        print(m, tag, 'differing', len(diff), 'idx', diff[:12], 'max abs', np.abs(v.logit.values - base.logit.values).max())
