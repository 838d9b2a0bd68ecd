"""Generate tests/golden/reference_objective.npz from the UNMODIFIED reference.

Run in the build container only (imports /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_objective.py

What the reference can provide for the objective, probed here and recorded in the fixture:
  * loss values: the jitclass singletons of loss.py:74-80 (`loss(p, y)`), on a grid that
    covers the logistic +-18 clamps and the squared-hinge kink;
  * Omega: ONLY `OmegaCS._eval` (omegacs.py:22-33) executes.  Every other `eval` of
    regularizer/*.py fails to type-check under numba (they call np.linalg.norm with
    signatures nopython mode does not support, or reshape with the axes swapped) -- the
    reference never calls them.  `eval_status` stores, per regularizer,
    whether the call succeeded, so the test documents exactly what is pinned.
The remaining Omega definitions are pinned in tests/test_oracle_golden.py against brute-force
formulas (itertools.combinations for the elementary symmetric polynomials).
"""
import json
import os
import sys

import numpy as np

REF = os.environ.get("SPARSEPOLY_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

from sparsepoly.loss import CLASSIFICATION_LOSSES  # noqa: E402
from sparsepoly.regularizer import REGULARIZATION  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden",
                   "reference_objective.npz")


def main():
    rng = np.random.RandomState(123)
    out = {}
    p = np.concatenate([rng.randn(40) * 3, [-30.0, -18.5, -18.0, 18.0, 18.5, 30.0, 1.0, -1.0, 0.0]])
    y = np.where(rng.rand(len(p)) < 0.5, -1.0, 1.0)
    out["loss_p"], out["loss_y"] = p, y
    for name in ("squared", "logistic", "squared_hinge"):
        obj = CLASSIFICATION_LOSSES[name]
        out[f"loss_{name}"] = np.array([obj.loss(float(a), float(b)) for a, b in zip(p, y)])
    status = {}
    cases = [(rng.randn(9, 3) * (rng.rand(9, 3) < 0.7), 2), (rng.randn(12, 4), 3), (rng.randn(30, 5), 4)]
    for t, (P_dk, deg) in enumerate(cases):
        out[f"P_{t}"] = P_dk
        out[f"deg_{t}"] = np.array(deg)
        out[f"omegacs_{t}"] = REGULARIZATION["omegacs"]()._eval(np.ascontiguousarray(P_dk[None]), deg)
    P_dk = cases[0][0]
    for name, R in REGULARIZATION.items():
        reg = R()
        for label, call in (("eval(P)", lambda: reg.eval(P_dk)), ("eval(P, 2)", lambda: reg.eval(P_dk, 2))):
            try:
                call()
                status[f"{name}.{label}"] = "ran"
            except Exception as e:  # numba TypingError / ValueError / NotImplementedError
                status[f"{name}.{label}"] = type(e).__name__
    out["eval_status"] = np.array(json.dumps(status, sort_keys=True))
    np.savez_compressed(OUT, **out)
    print(json.dumps(status, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
