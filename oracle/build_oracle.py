"""Builds the oracle's C restatement (oracle/_build/liboracle.so).  Checker / CPU baseline only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "liboracle.so")


def build(force=False):
    src = os.path.join(HERE, "corr_oracle.c")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(src):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-o", OUT, src, "-lm"], check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
