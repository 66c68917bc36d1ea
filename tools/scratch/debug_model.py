import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))  # repo root (run as `python tools/<name>.py`)

import sys; sys.path.insert(0, '.')
import numpy as np, torch
from tests import model_helpers as H
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
golden = H.load_golden(); cuda = torch.device('cuda:0')
for name in ['small', 'ref53', 'bl']:
    net = H.make_product(name, 'train', golden).to(cuda); net.train()
    data = H.make_data(name, cuda)
    ep = net(data)
    for k in H.EP_KEYS:
        want = golden[name + '_train_' + k]; got = ep[k].detach().cpu().numpy()
        print(name, 'train', k, got.dtype, 'maxdiff', np.abs(got.astype(np.float64) - want).max())
    loss = net.loss(ep, data)
    for k, v in loss.items():
        print(name, 'loss', k, v.item(), float(golden['%s_loss_%s' % (name, k)]))
    loss['total'].backward()
    params = dict(net.named_parameters())
    for key in [k for k in golden.files if k.startswith(name + '_grad_')]:
        pk = key[len(name) + 6:]; want = golden[key]; got = params[pk].grad.cpu().numpy()
        print(name, 'grad', pk, 'maxdiff', np.abs(got - want).max(), 'scale', np.abs(want).max())
    keys = list(golden['%s_gradnorm_keys' % name]); vals = golden['%s_gradnorm_vals' % name]
    bad = []
    for k, want in zip(keys, vals):
        g = params[str(k)].grad; got = g.double().norm().item() if g is not None else -1.0
        if abs(got - want) > 5e-3 * abs(want) + 1e-5: bad.append((str(k), got, want))
    print(name, 'gradnorm bad:', bad[:20], len(bad))
    # generate
    net = H.make_product(name, 'test', golden).to(cuda); net.eval()
    with torch.no_grad():
        ep, ed, parsed = net.generate(data)
    for k in H.EP_KEYS:
        want = golden[name + '_gen_' + k]; got = ep[k].cpu().numpy()
        print(name, 'gen', k, 'maxdiff', np.abs(got.astype(np.float64) - want).max())
    pm, wm = ed['pred_mask'], golden[name + '_gen_pred_mask']
    print(name, 'pred_mask diff at', np.argwhere(pm != wm).tolist(), 'got', np.argwhere(pm).tolist(), 'want', np.argwhere(wm).tolist())
    print(name, 'corners maxdiff', np.abs(parsed['pred_corners_3d'] - golden[name + '_gen_corners']).max())
    if (pm != wm).any():
        from oracle import geometry_ref as G
        hip = data['input_joints'][:, :, 0].cpu().numpy()
        r = G.parse_predictions(golden[name+'_gen_center'], golden[name+'_gen_size'], golden[name+'_gen_heading'], golden[name+'_gen_objectness_scores'], golden[name+'_gen_sem_cls_scores'], hip)
        print('oracle on golden outputs: mask==want', np.array_equal(r['pred_mask'], wm), 'nonempty sum', r['nonempty'].sum())
        from pose2room_b200 import geometry
        c, a, ne = geometry.decode_boxes(ep['center'], ep['size'], ep['heading'], data['input_joints'][:, :, 0])
        print('gpu nonempty vs oracle nonempty diff', np.argwhere(ne.cpu().numpy() != r['nonempty']).tolist())
        sc = parsed['obj_prob'][0]
        order = np.argsort(-sc)
        print('top scores', [(int(i), float(sc[i]), int(ne[0, i])) for i in order[:15]])
