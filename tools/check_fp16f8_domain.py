"""Counts hidden activations outside the validated fp16f8 domain (|a| >= 4096) during real renders, using a debug build of the
library (-DS2L_DBG_SATCOUNT).   python tools/check_fp16f8_domain.py [weight-scale ...]   (default scales: 1 2 4)"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util
_spec = importlib.util.spec_from_file_location("s2l_build", os.path.join(ROOT, "speech2lip_b200", "csrc", "build.py"))   # not via the package:
B = importlib.util.module_from_spec(_spec)                                                                                # it binds the library
_spec.loader.exec_module(B)
so = os.path.join(ROOT, "tools", "dbg_satcount.so")
if not os.path.exists(so):
    B.build(force=True, defines=("S2L_DBG_SATCOUNT",), out=so)
os.environ["S2L_LIB_PATH"] = so
import torch
import speech2lip_b200 as s2l
from speech2lip_b200 import _cabi, synth
lib = _cabi.lib()
for f in ("s2l_debug_sat_count_tc", "s2l_debug_sat_count_tc2"):
    getattr(lib, f).restype = C.c_ulonglong
dev = torch.device("cuda:0")
scales = [float(x) for x in sys.argv[1:]] or [1.0, 2.0, 4.0]
H = W = 64
for sc in scales:
    sd = synth.make_state_dict(0, "kaiming", 2, 3)
    sd = {k: (v * sc if k.startswith("pts_linears") and k.endswith("weight") else v) for k, v in sd.items()}
    w = s2l.PackedWeights({k: torch.from_numpy(v).to(dev) for k, v in sd.items()}, 2, 3)
    r = s2l.LipRenderer(w, "fp16f8")
    audio = torch.from_numpy(synth.make_audio(2, seed=3)).to(dev)
    lib.s2l_debug_sat_count_tc(); lib.s2l_debug_sat_count_tc2()
    out = r.render_frames(audio, torch.tensor([1, 2]), H, W, precision="fp16f8")
    exact = r.render_frames(audio, torch.tensor([1, 2]), H, W, precision="fp32")
    n = lib.s2l_debug_sat_count_tc() + lib.s2l_debug_sat_count_tc2()
    grid = torch.stack(torch.meshgrid(torch.linspace(0, 1, 32, device=dev), torch.linspace(0, 1, 32, device=dev), indexing="ij"), -1)
    rep = r.probe_fp16f8_domain(audio, torch.tensor([1]), grid.reshape(-1, 2))
    rel = ((out - exact).abs().max() / exact.abs().max()).item()
    print("hidden weights x%g: %d activations >= 4096 counted in the kernel, probe: max activation %.3g (ok=%s), max rel error vs fp32 %.2e, meta %s"
          % (sc, n, rep["max_activation"], rep["ok"], rel, w.meta()))
