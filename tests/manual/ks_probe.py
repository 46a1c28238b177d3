"""KS / t-test probe of whole runs, GPU against the oracle, over 120 seeds (run by hand under gpurun from the repo root)."""
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import pansim_b200 as pb
from pansim_b200 import simulate
from oracle import binding as ob
from scipy.stats import ks_2samp, ttest_ind
kw = dict(pop_size=64, core_size=20000, pan_genes=600, core_genes=200, n_gen=30, max_distances=2000)
S = 120
g, c = {}, {}
for seed in range(S):
    p = pb.Params(seed=seed, **kw); d = pb.derive(p)
    r = simulate.run(p, outpref=None)
    s = simulate.summarize(r.core_distances, r.acc_distances, r.gene_freqs, d.pan_size)
    o = ob.run(ob.default_params(seed=5000 + seed, threads=4, **kw))
    for k in s:
        g.setdefault(k, []).append(s[k]); c.setdefault(k, []).append(getattr(o, k))
for k in g:
    a, b = np.array(g[k]), np.array(c[k])
    if np.ptp(np.concatenate([a, b])) == 0: continue
    print(f"{k:18s} gpu {a.mean():.5f} +- {a.std()/np.sqrt(S):.5f}   cpu {b.mean():.5f} +- {b.std()/np.sqrt(S):.5f}   ks p={ks_2samp(a,b).pvalue:.4f} t p={ttest_ind(a,b).pvalue:.4f}")
