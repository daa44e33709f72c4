"""TEST INFRASTRUCTURE ONLY — import shim for the *real* reference implementation.

It imports the reference's own code from /root/reference (build container) or, where that does not exist (the GPU
box), from the unmodified copy oracle/build_ref.py vendored into oracle/_ref/ (git-ignored), to
(a) validate oracle/s2l_oracle.py against the reference's own code, (b) generate the golden vectors committed under
tests/golden/, (c) run the reference itself as bench.py's CPU / eager-GPU baselines and (d) drive the reference's own
callers against the drop-in in tests/.  It is never imported by the product package.

Recipe follows SURVEY.md §8(c): the reference's import chain pulls packages that
are not installed (lpips, imageio, librosa, flowlib->png, matplotlib) but that
none of the hot-path functions use, so they are replaced with empty stub modules.
"""
import os
import sys
import types

_VENDORED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REF_ROOT = os.environ.get("S2L_REFERENCE_ROOT") or ("/root/reference" if os.path.isdir("/root/reference/src") else _VENDORED)


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "face_simple", "models", "tf_nerf.py"))


def _install_stubs():
    for name in ("lpips", "imageio", "librosa", "librosa.filters", "flowlib"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        cm = types.ModuleType("matplotlib.cm")

        class _Cmap:
            N = 256

            def __call__(self, x=None, *a, **k):
                import numpy as _np
                n = 1 if x is None else _np.asarray(x).shape[0]
                return _np.zeros((n, 4))

        cm.get_cmap = lambda *a, **k: _Cmap()
        colors = types.ModuleType("matplotlib.colors")

        class _Listed:
            def __init__(self, *a, **k):
                pass

        class _Linear:
            @staticmethod
            def from_list(*a, **k):
                return _Cmap()

        colors.ListedColormap = _Listed
        colors.LinearSegmentedColormap = _Linear
        mpl.cm = cm
        mpl.colors = colors
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.cm"] = cm
        sys.modules["matplotlib.colors"] = colors


def load_reference():
    """Returns a namespace with the reference's TalkingFace, get_coords,
    density2outputs, get_rays, Trainer and the merged may.yaml config."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        from src import config as ref_config
        from src.face_simple.models.tf_nerf import TalkingFace
        from src.face_simple.rendering import get_coords, density2outputs
        from src.common import get_rays
        from src.face_simple.training import Trainer
        cfg = ref_config.load_config("configs/face_simple_configs/may/may.yaml", "configs/default.yaml")
    finally:
        os.chdir(cwd)
    # no dataset files here: fall back to the randn depth Parameter (tf_nerf.py:193-194)
    cfg["model"].pop("canonical_depth_init_path", None)
    ns = types.SimpleNamespace(TalkingFace=TalkingFace, get_coords=get_coords,
                               density2outputs=density2outputs, get_rays=get_rays,
                               Trainer=Trainer, cfg=cfg)
    return ns
