import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))  # repo root (run as `python tools/<name>.py`)

import sys; sys.path.insert(0, '.')
import numpy as np, torch
from tests import model_helpers as H
from oracle.model_ref import RefP2RNet
golden = H.load_golden()
for name in ['small', 'ref53']:
    B, T, J, S, P = H.CONFIGS[name]
    from pose2room_b200.p2rnet import P2RNet
    template = P2RNet(H.make_cfg(name, 'train')).state_dict()
    sd = H.weights_for(name, template, golden)
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    net = RefP2RNet(sd64, joint_num=J, num_seeds=S, num_target=P, training=True)
    data = H.make_data(name)
    d64 = {k: (v.double() if isinstance(v, torch.Tensor) and v.is_floating_point() else v) for k, v in data.items()}
    # keep index decisions identical to fp32: FPS/ball query run in float inside RefExt -> cast
    import oracle.model_ref as M
    orig_fps, orig_bq = M.RefExt.furthest_point_sampling, M.RefExt.ball_query
    M.RefExt.furthest_point_sampling = staticmethod(lambda x, m: orig_fps(x.float().contiguous(), m))
    M.RefExt.ball_query = staticmethod(lambda a, b, r, n: orig_bq(a.float().contiguous(), b.float().contiguous(), r, n))
    torch.set_default_dtype(torch.float64)
    ep = net.forward(d64)
    loss = net.loss(ep, d64)
    loss['total'].backward()
    torch.set_default_dtype(torch.float32)
    M.RefExt.furthest_point_sampling, M.RefExt.ball_query = orig_fps, orig_bq
    print(name, 'seed_inds same', np.array_equal(ep['seed_inds'].numpy(), golden[name+'_train_seed_inds']), 'fps same', np.array_equal(ep['aggregated_vote_inds'].numpy(), golden[name+'_train_aggregated_vote_inds']))
    for key in [k for k in golden.files if k.startswith(name + '_grad_')]:
        pk = key[len(name) + 6:]
        want = golden[key]; truth = net.p[pk].grad.numpy()
        print(name, pk, 'ref-fp32 vs fp64 truth: maxdiff %.3g scale %.3g rel %.2g' % (np.abs(want - truth).max(), np.abs(truth).max(), np.abs(want - truth).max() / np.abs(truth).max()))
    np.savez('scratch/fp64_grads_%s.npz' % name, **{k: net.p[k].grad.numpy() for k in net.p if net.p[k].grad is not None})
