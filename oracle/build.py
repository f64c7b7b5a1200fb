"""Build recipe for the CPU oracle (test infrastructure, never shipped on the product path).

    python oracle/build.py          -> oracle/libgs_oracle.so

There is no `oracle/_ref`: the reference's implementation of this path is the un-vendored pip
package diff-gaussian-rasterization==0.0.0 (/root/reference/environment.yml:129); its sources are
not under /root/reference, so there is nothing to compile from "where they lie" (DESIGN.md §3).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gs_oracle.c")
OUT = os.path.join(HERE, "libgs_oracle.so")


def build(force: bool = False) -> str:
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    cmd = [
        "gcc", "-O2", "-fPIC", "-shared", "-std=c11",
        # the canonical order is spelled with fmaf(); gcc must not add or remove fusions
        "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden",
        "-fopenmp", "-Wall", "-Wextra", "-o", OUT, SRC, "-lm",
    ]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
