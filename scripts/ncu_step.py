"""Tiny driver for `ncu --set full` captures of the SIMT / row kernels of the config-2 step at the sizes they run at
(B = 4096, D = 10, H = 600): five eager (no CUDA graph) steps of PvaeMnist through TrainStep.  Capture one step with
-k regex:<kernels> --launch-skip <launches of the first four steps> --launch-count <launches of one step>."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200")):
    sys.path.insert(0, p)
import torch
import hvae  # noqa: F401
from hvae import models, train

dev = torch.device("cuda")
torch.manual_seed(0)
x = torch.rand(4096, 1, 28, 28, device=dev)
model = models.PvaeMnist().to(dev)
ts = train.TrainStep(model, x, use_graph=False)
for _ in range(5):
    loss = ts.run()
torch.cuda.synchronize()
print("ok", float(loss if torch.is_tensor(loss) else 0.0))
