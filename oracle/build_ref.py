"""TEST / BASELINE INFRASTRUCTURE ONLY — vendors the UNMODIFIED reference into oracle/_ref/.

The reference (CVMI-Lab/Speech2Lip) is pure Python with no package metadata, so it cannot be pip-installed, and
/root/reference does not exist on the GPU box.  This recipe copies the handful of files its hot path needs
(src/**, configs/**, inference.py, train.py — ~200 KB, no modification) from where they lie under /root/reference into
oracle/_ref/, which is listed in .gitignore (reference sources never enter this repository's history) but NOT in
.gpurunignore, so the copy travels to the GPU box like a built .so.  There oracle/ref_shim.py imports it with the stub
modules of SURVEY.md §8(c), which lets
  * `bench.py --impl reference` time the reference's OWN code on the host cores (cpu_baseline.kind = "reference"),
  * bench.py's extras time the same module in PyTorch eager on the B200 ("the number a user of the reference sees"),
  * tests/test_gpu_real_callers.py run the reference's own callers (Trainer.predict_lip_image, the inference loop body)
    against the drop-in TalkingFace.
Run by __graft_entry__.build() whenever /root/reference is present:   python oracle/build_ref.py
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("S2L_REFERENCE_SRC", "/root/reference")
WANT = ["src", "configs", "inference.py", "train.py"]


def build(src=SRC, dst=DST):
    if not os.path.isfile(os.path.join(src, "src", "face_simple", "models", "tf_nerf.py")):
        return None
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    n = 0
    for item in WANT:
        s = os.path.join(src, item)
        if os.path.isdir(s):
            for root, dirs, files in os.walk(s):
                dirs[:] = [d for d in dirs if d != "__pycache__"]
                for f in files:
                    if f.endswith((".py", ".yaml", ".yml")):
                        rel = os.path.relpath(os.path.join(root, f), src)
                        os.makedirs(os.path.dirname(os.path.join(dst, rel)), exist_ok=True)
                        shutil.copyfile(os.path.join(root, f), os.path.join(dst, rel))
                        n += 1
        elif os.path.isfile(s):
            shutil.copyfile(s, os.path.join(dst, item))
            n += 1
    with open(os.path.join(dst, "VENDORED_FROM"), "w") as fh:
        fh.write("%s (unmodified copy made by oracle/build_ref.py; %d files)\n" % (src, n))
    return dst


if __name__ == "__main__":
    out = build()
    print(out or "reference tree not present at %s — nothing vendored" % SRC)
    sys.exit(0)
