"""CPU: oracle/model_ref.py (the plain-PyTorch port used as CPU baseline and referee on the GPU box) against
goldens produced by the unmodified reference model (tests/golden/make_golden_model.py)."""
import numpy as np
import pytest
import torch

from oracle.model_ref import RefP2RNet
from tests import model_helpers as H


@pytest.fixture(scope="module")
def golden():
    return H.load_golden()


def _oracle(name, golden, training):
    B, T, J, S, P = H.CONFIGS[name]
    from pose2room_b200.p2rnet import P2RNet
    template = P2RNet(H.make_cfg(name, "train")).state_dict()
    return RefP2RNet(H.weights_for(name, template, golden), joint_num=J, num_seeds=S, num_target=P, training=training)


@pytest.mark.parametrize("name", ["small", "ref53", "bl"])
def test_oracle_train_forward_loss_backward_vs_reference(golden, name):
    net = _oracle(name, golden, training=True)
    data = H.make_data(name)
    ep = net.forward(data)
    for k in H.EP_KEYS:
        want = golden["%s_train_%s" % (name, k)]
        got = ep[k].detach().numpy()
        if want.dtype.kind in "iu":
            assert np.array_equal(got, want), k
        else:
            assert np.allclose(got, want, rtol=1e-5, atol=1e-6), (k, np.abs(got - want).max())
    loss = net.loss(ep, data)
    for k, v in loss.items():
        assert abs(v.item() - float(golden["%s_loss_%s" % (name, k)])) < 1e-5 * max(1.0, abs(v.item())), k
    if name == "bl":
        return
    loss["total"].backward()
    for key in [k for k in golden.files if k.startswith(name + "_grad_")]:
        pk = key[len(name) + 6:]
        want = golden[key]
        got = net.p[pk].grad.numpy()
        assert np.allclose(got, want, rtol=2e-4, atol=1e-6 + 2e-4 * np.abs(want).max()), pk


@pytest.mark.parametrize("name", ["small", "ref53", "bl"])
def test_oracle_generate_vs_reference(golden, name):
    net = _oracle(name, golden, training=False)
    ep, parsed = net.generate(H.make_data(name))
    for k in H.EP_KEYS:
        want = golden["%s_gen_%s" % (name, k)]
        got = ep[k].numpy()
        if want.dtype.kind in "iu":
            assert np.array_equal(got, want), k
        else:
            assert np.allclose(got, want, rtol=1e-5, atol=1e-6), (k, np.abs(got - want).max())
    assert np.array_equal(parsed["pred_mask"], golden["%s_gen_pred_mask" % name])
    assert np.allclose(parsed["corners"], golden["%s_gen_corners" % name], atol=1e-5)
