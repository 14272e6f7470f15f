import os, sys, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import gpuaudiobench_b200 as g
from oracle_lib import Oracle
o = Oracle()
def snr(a, b):
    b = b.astype(np.float64); return 10*np.log10((b**2).sum()/max(((a.astype(np.float64)-b)**2).sum(), 1e-300))
T, B, L = 128, 512, 16384
x = o.generate_input(T*B); h = o.generate_ir(T, L, "direct")
padded = np.concatenate([np.zeros(L-1, np.float32), x])
hist = np.stack([padded[t*B:t*B+L-1] for t in range(T)])
r1 = o.r1(x, h, L, B, T)
for rep in range(4):
    for algo in (g.ALGO_DIRECT, g.ALGO_UPOLS, g.ALGO_DIRECT_TC):
        with g.ConvEngine(T, B, L, algo) as e:
            e.load_ir(h); e.prime_history(hist)
            y, _ = e.process_host(x.reshape(T, B))
        s = np.array([snr(y[t], r1[t]) for t in range(T)])
        bad = np.where(s < 100)[0]
        print(rep, algo, "min %.1f at %d; tracks below 100 dB: %s" % (s.min(), s.argmin(), bad.tolist()[:40]), flush=True)
