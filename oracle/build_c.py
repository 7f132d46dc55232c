"""Builds oracle/_cbuild/libsipref.so from oracle/c/ref_kernels.c (gcc + OpenMP).

TEST / MEASUREMENT INFRASTRUCTURE ONLY: the threaded CPU restatement timed by bench.py's cpu_baseline and
`--impl reference` legs.  No -march flag: the library is built in one container and run on another host."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "ref_kernels.c")
DEPS = [SRC, os.path.join(HERE, "c", "ref_kernels_body.inc"), os.path.join(HERE, "c", "ref_kernels.h")]
OUT = os.path.join(HERE, "_cbuild", "libsipref.so")


def needs_build() -> bool:
    return not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS)


def build(force: bool = False) -> str:
    if force or needs_build():
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run(["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-Wall", "-o", OUT, SRC, "-lm"],
                       check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
