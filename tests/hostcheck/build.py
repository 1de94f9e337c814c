"""TEST INFRASTRUCTURE: build tests/hostcheck/libazg_hostcheck.so (x86 build of the arena core).
-ffp-contract=off: the core spells out every rounding; the host compiler must not fuse."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "libazg_hostcheck.so")
SRC = os.path.join(HERE, "arena_host.cpp")
DEPS = [SRC] + [os.path.join(HERE, "..", "..", "alphazero-gnn_b200", "csrc", f)
                for f in ("azg_arena_core.cuh", "azg_rules.cuh", "azg_common.cuh")]


def build():
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O1", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
           SRC, "-o", OUT]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("hostcheck build failed:\n" + r.stdout)
    return OUT


if __name__ == "__main__":
    print(build())
