import sys, torch
from pose2room_b200 import _lib
dev = torch.device("cuda:0")
M, C, P = int(sys.argv[1]), 64, 25
g = torch.Generator().manual_seed(1)
x = torch.randn(M, C, generator=g).to(dev).bfloat16(); dy = torch.randn(M, C, generator=g).to(dev).bfloat16()
st = (torch.rand(4, C, generator=g) + 0.5).to(dev)
s = torch.zeros(2, C, dtype=torch.float64, device=dev)
dx = torch.empty_like(x); cs = torch.zeros(P, C, dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for it in range(3):
    _lib.call("p2r_bn_bwd_apply", dy.data_ptr(), x.data_ptr(), None, 1, M, C, st[0].data_ptr(), st[1].data_ptr(),
              st[2].data_ptr(), s[0].data_ptr(), s[1].data_ptr(), 2, dx.data_ptr(), None, st[3].data_ptr(), cs.data_ptr(), P, stream)
    torch.cuda.synchronize()
    print("iter", it, "ok", float(cs.sum()), flush=True)
