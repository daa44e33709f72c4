"""Builds speech2lip_b200/csrc/libs2l_b200.so with nvcc for sm_100a (in-tree, no JIT cache)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["s2l_capi.cu", "s2l_pack.cu", "s2l_audio.cu", "s2l_mlp_fp32.cu", "s2l_mlp_tc.cu", "s2l_mlp_tc2.cu", "s2l_reduce.cu", "s2l_postfusion.cu", "s2l_mlp_bwd.cu"]
LIB = os.path.join(HERE, "libs2l_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "--shared", "-cudart", "static"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(HERE, "..", "..", "include", "speech2lip_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    out = out or LIB
    if not force and not needs_build() and out == LIB:
        return LIB
    cmd = [NVCC] + FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) \
        + [os.path.join(HERE, s) for s in SOURCES] + ["-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building %s" % out)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
