"""Build the oracle's C restatement (oracle/nr_raster.c, oracle/fma_mm.c) into oracle/libnr_oracle.so.

TEST INFRASTRUCTURE.  -ffp-contract=off is part of the oracle's arithmetic contract (no FMA).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRCS = [os.path.join(HERE, "nr_raster.c"), os.path.join(HERE, "fma_mm.c")]
OUT = os.path.join(HERE, "libnr_oracle.so")


def build(force=False):
    if (not force and os.path.exists(OUT)
            and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in SRCS)):
        return OUT
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared",
           "-fvisibility=hidden", "-Wall", "-o", OUT] + SRCS + ["-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
