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
P = dict(net.named_parameters())
for k in ['backbone.edge_importance.3', 'backbone.st_gcn_networks.3.gcn.conv.weight', 'backbone.st_gcn_networks.3.tcn.0.bias', 'backbone.st_gcn_networks.3.tcn.0.weight', 'backbone.st_gcn_networks.3.tcn.3.bias', 'backbone.st_gcn_networks.3.tcn.3.weight','backbone.st_gcn_networks.3.tcn.2.weight']:
    t = truth[k]; g = P[k].grad.double().cpu().numpy()
    d = np.abs(g - t)
    idx = np.argsort(-d.ravel())[:6]
    print(k, t.shape, 'max', d.max(), 'mean', d.mean(), 'scale', np.abs(t).max())
    for i in idx:
        ui = np.unravel_index(i, t.shape)
        print('   ', ui, g[ui], t[ui])
# run the whole thing twice: is my result deterministic?
net.zero_grad(); ep = net(data); loss = net.loss(ep, data); loss['total'].backward()
g2 = P['backbone.edge_importance.3'].grad.double().cpu().numpy()
net.zero_grad(); ep = net(data); loss = net.loss(ep, data); loss['total'].backward()
g3 = P['backbone.edge_importance.3'].grad.double().cpu().numpy()
print('run-to-run diff', np.abs(g2 - g3).max())
