"""Runs bench.py's config-5 training-step measurement alone (tensor-core training path vs the reference's eager autograd).
    python tools/bench_train_tc.py [H W]"""
import importlib.util, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
spec = importlib.util.spec_from_file_location("s2l_bench", os.path.join(ROOT, "bench.py"))
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)
sizes = ((int(sys.argv[1]), int(sys.argv[2])),) if len(sys.argv) > 2 else ((80, 120), (256, 256))
torch.cuda.set_device(0)
print(json.dumps(b.train_step_extras(torch.device("cuda:0"), sizes=sizes), indent=1))
