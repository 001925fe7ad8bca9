import sys, time, torch, os
sys.path.insert(0, ".")
import torch.nn.functional as F
from mslesions3d_b200 import ops, _lib
torch.manual_seed(0)
cin, size, sd = int(sys.argv[1]), tuple(int(v) for v in sys.argv[2:5]), int(sys.argv[5])
x = torch.randn((1, cin) + size).to(torch.bfloat16)
w = torch.randn(32, cin, 3, 3, 3) * 0.2
scale, shift = torch.ones(32).cuda(), torch.zeros(32).cuda()
xc = x.cuda()
wp = ops.pack_stem_weight(w.cuda())
torch.cuda.synchronize(); print("inputs ready", flush=True)
lib = _lib.load()
n, c, d, h, ww = xc.shape
y = torch.empty((n, (d-1)//sd+1, (h-1)//2+1, (ww-1)//2+1, 32), dtype=torch.bfloat16, device="cuda")
for name in sys.argv[6:] or ["ssd3d_stem_conv_bn_relu"]:
    fn = getattr(lib, name)
    t0 = time.time()
    rc = fn(xc.data_ptr(), 1, wp.data_ptr(), scale.data_ptr(), shift.data_ptr(), y.data_ptr(), n, c, d, h, ww, sd, torch.cuda.current_stream().cuda_stream)
    try:
        torch.cuda.synchronize()
        print(name, "rc", rc, "ok in %.3fs" % (time.time() - t0), flush=True)
        want = F.relu(F.conv3d(x.float(), w.to(torch.bfloat16).float(), None, (sd, 2, 2), 1)).permute(0, 2, 3, 4, 1)
        print(" max err", float((y.float().cpu() - want).abs().max()), "ref max", float(want.abs().max()), flush=True)
    except Exception as e:
        print(name, "rc", rc, "failed after %.3fs: %s" % (time.time() - t0, str(e)[:120]), flush=True)
        break
