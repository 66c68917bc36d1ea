"""The three temporal-conv GEMMs (forward, input gradient, weight gradient) at the BASELINE shape, timed alone with CUDA
events; run once per setting of P2R_TCONV_HALO (read once per process).  Also checks them against nn.Conv2d on integers."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

from pose2room_b200 import _lib, gemm_sm100

dev = torch.device("cuda:0")
gemm_sm100.install()
B, T, V, C = 32, 1024, 25, 64
g = torch.Generator().manual_seed(1)
x = torch.randint(-2, 3, (B, T, V, C), generator=g).float().to(dev).bfloat16()
dy = torch.randint(-2, 3, (B * T * V, C), generator=g).float().to(dev).bfloat16()
conv = nn.Conv2d(C, C, (3, 1), (1, 1), (1, 0)).to(dev)
with torch.no_grad():
    conv.weight.copy_(torch.randint(-2, 3, conv.weight.shape, generator=g).float())
    conv.bias.copy_(torch.randint(-2, 3, (C,), generator=g).float())
w = conv.weight.detach()
w2 = w.bfloat16()[:, :, :, 0].permute(0, 2, 1).reshape(C, 3 * C).contiguous()
wt = w.bfloat16()[:, :, :, 0].permute(2, 0, 1).reshape(3 * C, C).contiguous()
bias = conv.bias.detach().float().contiguous()
y = torch.empty(B * T * V, C, dtype=torch.bfloat16, device=dev)
dx = torch.empty(B, T, V, C, dtype=torch.bfloat16, device=dev)
dw = torch.zeros(C, 3 * C, dtype=torch.float32, device=dev)
sums = torch.zeros(16, 2, 64, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
fwd = lambda: _lib.call("p2r_tconv_bf16", 0, x.data_ptr(), w2.data_ptr(), None, y.data_ptr(), B, T * V, C, C, 3, V, bias.data_ptr(), 1, sums.data_ptr(), 16, st)
bwd = lambda: _lib.call("p2r_tconv_bf16", 1, dy.data_ptr(), wt.data_ptr(), None, dx.data_ptr(), B, T * V, C, C, 3, V, None, 1, None, 1, st)
wgt = lambda: _lib.call("p2r_tconv_bf16", 2, x.data_ptr(), None, dy.data_ptr(), dw.data_ptr(), B, T * V, C, C, 3, V, None, 148, None, 1, st)


def t(f, n=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


fwd(); bwd(); dw.zero_(); wgt()
torch.cuda.synchronize()
xr = x.float().requires_grad_(True)
ref = conv(xr.permute(0, 3, 1, 2)).permute(0, 2, 3, 1).reshape(B * T * V, C)
rx, rw = torch.autograd.grad(ref, [xr, conv.weight], dy.float())
ok = (bool(torch.equal(y.float(), ref.detach())), bool(torch.equal(dx.float(), rx)),
      bool(torch.allclose(dw.reshape(C, 3, C).permute(0, 2, 1), rw[:, :, :, 0], rtol=1e-5, atol=1e-2)))
print("P2R_TCONV_HALO=%s exact fwd/dx/dw: %s   us: fwd %.1f dx %.1f dw %.1f" % (os.environ.get("P2R_TCONV_HALO", "(default)"), ok, t(fwd), t(bwd), t(wgt)), flush=True)

if os.environ.get("P2R_TCONV_TRACE"):
    import numpy as np
    for name, f in (("fwd", fwd), ("dx", bwd)):
        buf = torch.zeros(3, 64, 8, dtype=torch.int64, device=dev)
        _lib.call("p2r_debug_tconv_trace", buf.data_ptr())
        f()
        torch.cuda.synchronize()
        _lib.call("p2r_debug_tconv_trace", None)
        a = buf.cpu().numpy()
        t0 = a[a > 0].min()
        rel = np.where(a > 0, a - t0, -1)
        print(name, "producer (after empty wait, after issue) | MMA (after tmem_empty, after full, after commit) | epilogue (start, bias bar, tmem_full, pairA, staged, stored, stats)")
        for t in range(24):
            print("  tile %2d  P %s | M %s | E %s" % (t, rel[0, t, :2].tolist(), rel[1, t, :3].tolist(), rel[2, t, :7].tolist()))
