"""Builds speech2lip_b200/csrc/libs2l_b200.so with nvcc for sm_100a (in-tree, no JIT cache).
Translation units compile in parallel into build/*.o (only the stale ones), then link into one shared object."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["s2l_capi.cu", "s2l_pack.cu", "s2l_audio.cu", "s2l_mlp_fp32.cu", "s2l_mlp_tc.cu", "s2l_mlp_tc2.cu", "s2l_reduce.cu",
           "s2l_postfusion.cu", "s2l_mlp_bwd.cu", "s2l_train_dgrad.cu", "s2l_train_wgrad.cu", "s2l_train_final.cu", "s2l_gemm_fp32.cu", "s2l_peer.cu"]
LIB = os.path.join(HERE, "libs2l_b200.so")
OBJ_DIR = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
          "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]
LFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-cudart", "static"]


def _headers():
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cuh")]
    deps.append(os.path.join(HERE, "..", "..", "include", "speech2lip_b200.h"))
    return deps


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, f) for f in SOURCES] + _headers()
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, obj, defines, verbose):
    cmd = [NVCC] + CFLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False, defines=(), out=None):
    out = out or LIB
    if not force and not needs_build() and out == LIB:
        return LIB
    tag = "".join(sorted(defines))
    odir = os.path.join(OBJ_DIR, tag) if tag else OBJ_DIR
    os.makedirs(odir, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(HERE, s), os.path.join(odir, s[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((src, obj))
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4) or 1) as ex:
        results = list(ex.map(lambda j: _compile(j[0], j[1], defines, verbose), jobs))
    for (src, _), (rc, log) in zip(jobs, results):
        if verbose or rc != 0:
            sys.stderr.write(log)
        if rc != 0:
            raise RuntimeError("nvcc failed compiling %s" % src)
    cmd = [NVCC] + LFLAGS + [os.path.join(odir, s[:-3] + ".o") for s in SOURCES] + ["-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking %s" % out)
    return out


if __name__ == "__main__":
    print(build(force="-f" in sys.argv or "--force" in sys.argv or len(sys.argv) == 1, verbose="-v" in sys.argv))
