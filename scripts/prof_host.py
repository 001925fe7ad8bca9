"""cProfile of the host side of LSSD3D.predict_batches at the benchmark shape (is the pipelined step host-bound?)."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mslesions3d_b200.ssd3d import LSSD3D  # noqa: E402
from mslesions3d_b200 import synthetic as O  # noqa: E402

dev = torch.device("cuda")
model = LSSD3D(n_classes=2, input_channels=2, input_size=(128, 128, 128))
model.load_state_dict(O.random_state_dict(2, seed=0))
model = model.to(dev).eval()
xs = [torch.randn(8, 2, 128, 128, 128, device=dev).to(torch.bfloat16) for _ in range(4)]
steps = 300
with torch.no_grad():
    for _ in model.predict_batches({"img": xs[i % 4]} for i in range(12)):
        pass
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in model.predict_batches({"img": xs[i % 4]} for i in range(steps)):
        pass
    torch.cuda.synchronize()
    print("wall per step: %.1f us" % ((time.perf_counter() - t0) / steps * 1e6))
    pr = cProfile.Profile()
    pr.enable()
    for _ in model.predict_batches({"img": xs[i % 4]} for i in range(steps)):
        pass
    pr.disable()
    torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
