import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))  # repo root (run as `python tools/<name>.py`)

import sys, traceback; sys.path.insert(0, '.')
import numpy as np, torch
from pose2room_b200 import gemm_sm100, ops, synthetic
from pose2room_b200.config import P2RConfig
from pose2room_b200.p2rnet import P2RNet
dev = torch.device('cuda:0'); gemm_sm100.install()
torch.manual_seed(42); np.random.seed(42)
net = P2RNet(P2RConfig(mode='train', joint_num=25, num_frames=1024, precision='bf16'))
net.load_state_dict(synthetic.deterministic_state_dict(net.state_dict(), seed=7)); net = net.to(dev).train()
params = [p for p in net.parameters() if p.requires_grad]
opt = torch.optim.AdamW(params, lr=1e-3, fused=True, capturable=True)
data = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synthetic.make_batch(4, 1024, 25, seed=1).items()}
for it in range(3):
    try:
        opt.zero_grad(set_to_none=True)
        with ops.overlap_weight_grads():
            ep = net(data); loss = net.loss(ep, data)['total']
            loss.backward()
        opt.step(); torch.cuda.synchronize(); print('step', it, 'ok', float(loss))
    except Exception:
        traceback.print_exc(); break
