import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from ldsr_b200 import _lib
sys.argv = ["x"]
exec(open("tools/profile_scan.py").read().split("T = int")[0])
T=100000; d,p,q=4,20,20
rng = np.random.default_rng(1)
u = rng.standard_normal((p, T)); v = rng.standard_normal((q, T)); y = rng.standard_normal(T); y[rng.uniform(size=T) < 0.1] = np.nan
th = np.stack([model(rng, d, p, q, T) for _ in range(int(os.environ.get("NF","1")))])
for i in range(2):
    r=_lib.smoother_d(d, y, u, v, th, method=1, want=())
print(r["kernel_ms"])
