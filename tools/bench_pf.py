"""runs bench.py's post-fusion extra alone"""
import importlib.util, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
spec = importlib.util.spec_from_file_location("s2l_bench", os.path.join(ROOT, "bench.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
torch.cuda.set_device(0)
print(json.dumps(b.post_fusion_extras(torch.device("cuda:0")), indent=1))
