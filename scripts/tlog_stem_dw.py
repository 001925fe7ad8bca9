import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
os.environ["SSD3D_SDW_TLOG"] = str(buf.data_ptr())
from mslesions3d_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((8, 2, 128, 128, 128), device="cuda", generator=g).to(torch.bfloat16)
ws = ops.pack_stem_weight(torch.randn((32, 2, 3, 3, 3), device="cuda", generator=g) * 0.2)
wd = ops.pack_dw_weight(torch.randn((32, 1, 3, 3, 3), device="cuda", generator=g) * 0.3)
sc0, sh0 = torch.rand(32, device="cuda") + 0.5, torch.randn(32, device="cuda") * 0.1
for _ in range(3):
    y = ops.stem_dw_bn_relu(x, ws, sc0, sh0, wd, sc0, sh0, 2)
torch.cuda.synchronize()
allb = buf.cpu()
t = allb[:384].view(3, 32, 4)
t0 = int(t[t > 0].min())
names = ["producer: job issued", "mma: start | acc0 free | acc1 free | done", "compute: start | A done | bar | B done"]
for r in range(3):
    print(names[r])
    for j in range(22):
        row = [int(v) - t0 if int(v) > 0 else -1 for v in t[r, j]]
        print("  job %2d: %s" % (j, "  ".join("%7d" % v for v in row)))

uu = allb[512:512 + 72].view(4, 2, 3, 3)
print("mma units of jobs 4..7: (c, kd): wait start | wait done | issued+committed   (relative to job's first stamp)")
for jj in range(4):
    base = int(uu[jj, 0, 0, 0])
    print("  job %d: " % (jj + 4) + " | ".join("%d,%d: %5d %5d %5d" % (c, kd, int(uu[jj, c, kd, 0]) - base, int(uu[jj, c, kd, 1]) - base, int(uu[jj, c, kd, 2]) - base) for c in range(2) for kd in range(3)))
