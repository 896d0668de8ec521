"""Builds the oracle's C restatement (oracle/_build/liboracle.so).  Checker / CPU baseline only."""
import os
import subprocess

import hashlib

HERE = os.path.dirname(os.path.abspath(__file__))


def _cpu_tag():
    """-march=native code only runs on CPUs with the same ISA extensions: key the file on the CPU flags, so that a library
    built in one container is rebuilt (gcc is in the image) rather than loaded on a box with a different CPU."""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return hashlib.md5(" ".join(sorted(line.split(":", 1)[1].split())).encode()).hexdigest()[:10]
    except OSError:
        pass
    return "generic"


OUT = os.path.join(HERE, "_build", "liboracle-%s.so" % _cpu_tag())


def build(force=False):
    src = os.path.join(HERE, "corr_oracle.c")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(src):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-o", OUT, src, "-lm"], check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
