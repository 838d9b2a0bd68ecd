import os, sys, warnings, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
from sparsepoly_b200 import synth
from oracle import oracle as O
import sparsepoly_b200 as S
n,d,r=20000,2000,20
X = synth.uniform_sparse(n, d, r, 1)
rng=np.random.RandomState(3)
beta_=np.zeros(d); act=rng.choice(d,d//10,replace=False); beta_[act]=rng.randn(act.size)
s=X@beta_; s=s+0.5*(s**2-np.mean(s**2))+0.1*np.std(s)*rng.randn(n)
y=np.where(s>np.median(s),1.0,-1.0)
base=dict(degree=3, loss="logistic", n_components=2, solver="pcd", regularizer="omegati", alpha=1e-6, beta=1e-6, gamma=5e-10, mean=True, fit_linear=True, fit_lower="explicit", shuffle=False, random_state=0, tol=-1.0, max_iter=1)
variants = {"base": {}, "no_lower": dict(fit_lower=None), "l1": dict(regularizer="l1"), "sqhinge": dict(loss="squared_hinge"),
            "k1": dict(n_components=1), "nolinear": dict(fit_linear=False), "deg4": dict(degree=4), "gamma1e-8": dict(gamma=1e-8)}
for vname, v in variants.items():
    kw=dict(base, **v)
    out=O.fit_fm(X,y,**kw)
    for sweep in ("cluster","window"):
        os.environ["SPARSEPOLY_B200_SWEEP"]=sweep
        est=S.SparseFactorizationMachineClassifier(**kw).fit(X,y)
        errs=[float(np.abs(est.P_[o]-out["P_"][o]).max()) for o in range(est.P_.shape[0])]
        print(vname, sweep, est._dev_state["plan"].mode, "max abs diff per order", errs, "w", float(np.abs(est.w_-out["w_"]).max()), "max|P|", [float(np.abs(out["P_"][o]).max()) for o in range(est.P_.shape[0])], flush=True)
