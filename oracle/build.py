"""Build the C parity oracle (TEST INFRASTRUCTURE ONLY -- see sp_oracle.c header).

    python oracle/build.py

Produces oracle/libsp_oracle.so (git-ignored, travels to the GPU box with the snapshot).
-ffp-contract=off keeps gcc from fusing a*b+c, matching numba's codegen for the reference.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "sp_oracle.c")
OUT = os.path.join(HERE, "libsp_oracle.so")


def build(force=False):
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
           "-std=c11", "-Wall", "-Wextra", "-o", OUT, SRC, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
