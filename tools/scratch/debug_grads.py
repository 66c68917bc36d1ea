import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))  # repo root (run as `python tools/<name>.py`)

import sys; sys.path.insert(0, '.')
import numpy as np, torch
from tests import model_helpers as H
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
golden = H.load_golden(); cuda = torch.device('cuda:0')
name = 'small'
truth = np.load('scratch/fp64_grads_small.npz')
net = H.make_product(name, 'train', golden).to(cuda); net.train()
data = H.make_data(name, cuda)
ep = net(data); loss = net.loss(ep, data); loss['total'].backward()
rows = []
for k, p in net.named_parameters():
    if p.grad is None or k not in truth.files: continue
    t = truth[k]; g = p.grad.double().cpu().numpy()
    rows.append((np.abs(g - t).max() / (np.abs(t).max() + 1e-12), k, np.abs(t).max()))
rows.sort(reverse=True)
for r in rows[:25]: print('%.3g  %s  scale %.3g' % r)
print('...')
for r in rows[-5:]: print('%.3g  %s  scale %.3g' % r)
