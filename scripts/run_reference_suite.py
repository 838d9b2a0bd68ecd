"""Run the reference's OWN test-suite (tests/test_{pcd,pbcd,psgd,prox}.py, 864 differential tests against its
brute-force "slow" solvers) against this backend: `sparsepoly` is aliased to
`sparsepoly_b200` by a generated shim package.

    build container:  python scripts/run_reference_suite.py --stage      (copies the tests, git-ignored)
    GPU box:          python scripts/run_reference_suite.py [pytest args]

The reference's test files are copied to oracle/_ref/reference_tests/ (git-ignored, travels with gpurun like the
other oracle/_ref artefacts) -- they are never committed."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "oracle", "_ref", "reference_tests")
REF = os.environ.get("SPARSEPOLY_REFERENCE", "/root/reference")

SHIM = {
    "__init__.py": "from sparsepoly_b200 import (SparseAllSubsetsClassifier, SparseAllSubsetsRegressor,\n"
                   "    SparseFactorizationMachineClassifier, SparseFactorizationMachineRegressor)\n",
    "kernels.py": "from sparsepoly_b200.kernels import *  # noqa: F401,F403\n"
                  "from sparsepoly_b200.kernels import all_subsets_kernel, anova_kernel, poly_predict  # noqa: F401\n",
    "regularizer.py": "from sparsepoly_b200.regularizer import *  # noqa: F401,F403\n"
                      "from sparsepoly_b200.regularizer import L1, L21, SquaredL12, SquaredL21  # noqa: F401\n",
}


def stage():
    if not os.path.isdir(os.path.join(REF, "tests")):
        raise SystemExit(f"{REF}/tests not found (stage in the build container)")
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    shutil.copytree(os.path.join(REF, "tests"), os.path.join(DEST, "tests"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    os.makedirs(os.path.join(DEST, "sparsepoly"))
    for name, body in SHIM.items():
        with open(os.path.join(DEST, "sparsepoly", name), "w") as f:
            f.write(body)
    # the unmodified reference package itself, for bench.py --impl reference (numba is part of the image): the
    # reference arm then times the REAL reference on the GPU box's host cores instead of the C port
    pkg = os.path.join(ROOT, "oracle", "_ref", "reference_pkg")
    if os.path.isdir(pkg):
        shutil.rmtree(pkg)
    shutil.copytree(os.path.join(REF, "sparsepoly"), os.path.join(pkg, "sparsepoly"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    print("staged", DEST, "and", pkg)


def main():
    if "--stage" in sys.argv:
        stage()
        return 0
    if not os.path.isdir(os.path.join(DEST, "tests")):
        raise SystemExit("run with --stage in the build container first")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([DEST, ROOT, os.environ.get("PYTHONPATH", "")]),
               PYTHONDONTWRITEBYTECODE="1")
    args = sys.argv[1:] or ["-q", "-x", "--no-header", "-p", "no:cacheprovider"]
    return subprocess.call([sys.executable, "-m", "pytest", "tests"] + args, cwd=DEST, env=env)


if __name__ == "__main__":
    sys.exit(main())
