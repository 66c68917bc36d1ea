"""GPU parity of the whole hot path: pose2room_b200.p2rnet.P2RNet (fp32 parity mode) against goldens produced
by the UNMODIFIED reference model on the same seeded inputs and weights.

Tolerances (north_star): indices (seed_inds, FPS picks, NMS mask) bit-exact; centre / size / heading and the
other float outputs within 1e-4 absolute; losses within 1e-4 relative.  Gradients: relative L2 error below
1e-2 and 99 % of the entries within 2e-3 of the tensor's largest entry.  (Measured against an fp64 run, the
reference's own fp32 gradients carry isolated errors up to 1e-3 of scale: a pre-activation that lands within
1e-6 of zero flips its ReLU mask when the summation order changes -- one fused GEMM here vs conv + einsum
there -- and moves single rows of dW.  The bulk error is ~1e-6; see DESIGN.md "gradient parity".)"""
import numpy as np
import pytest
import torch

from tests import model_helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    return H.load_golden()


def _check_endpoints(ep, golden, prefix):
    for k in H.EP_KEYS:
        want = golden[prefix + k]
        got = ep[k].detach().cpu().numpy()
        if want.dtype.kind in "iu":
            assert np.array_equal(got, want), k
        else:
            assert got.dtype == want.dtype, (k, got.dtype, want.dtype)
            # 1e-4 absolute on the box tensors (north_star); the fixture's logits are deliberately 8x wider
            # (they reach |20|) and their fp32 noise scales with them: 2e-5 of the tensor's range there
            tol = max(1e-4, 2e-5 * float(np.abs(want).max())) if k.endswith("_scores") else 1e-4
            assert (np.abs(got - want) <= tol).all(), (k, np.abs(got - want).max())


@pytest.mark.parametrize("name", ["small", "ref53", "bl"])
def test_train_forward_loss_backward(cuda, golden, name):
    net = H.make_product(name, "train", golden).to(cuda)
    net.train()
    data = H.make_data(name, cuda)
    ep = net(data)
    _check_endpoints(ep, golden, name + "_train_")
    for k in ["seed_features", "vote_features"]:
        head = ep[k].detach()[:, :4, :32].cpu().numpy()
        assert np.abs(head - golden["%s_train_%s_head" % (name, k)]).max() < 1e-4, k
        t = ep[k].detach().double()
        stats = np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()])
        assert np.allclose(stats, golden["%s_train_%s_stats" % (name, k)], rtol=1e-4, atol=1e-2), k
    loss = net.loss(ep, data)
    assert set(loss) == {"total", "vote_loss", "objectness_loss", "center_loss", "size_loss", "heading_loss",
                         "sem_cls_loss", "pos_ratio", "neg_ratio", "obj_acc"}
    for k, v in loss.items():
        want = float(golden["%s_loss_%s" % (name, k)])
        assert abs(v.item() - want) < 1e-4 * max(1.0, abs(want)), (k, v.item(), want)
    loss["total"].backward()
    params = dict(net.named_parameters())
    for key in [k for k in golden.files if k.startswith(name + "_grad_")]:
        pk = key[len(name) + 6:]
        want = golden[key]
        got = params[pk].grad.cpu().numpy()
        diff = np.abs(got.astype(np.float64) - want)
        scale = np.abs(want).max()
        assert np.linalg.norm(diff) <= 1e-2 * np.linalg.norm(want) + 1e-7, (pk, np.linalg.norm(diff), np.linalg.norm(want))
        assert np.percentile(diff, 99) <= 2e-3 * scale + 1e-7, (pk, np.percentile(diff, 99), scale)
    keys = list(golden["%s_gradnorm_keys" % name])
    vals = golden["%s_gradnorm_vals" % name]
    for k, want in zip(keys, vals):
        k = str(k)
        g = params[k].grad
        got = g.double().norm().item() if g is not None else -1.0
        assert abs(got - want) <= 1e-2 * abs(want) + 1e-4, (k, got, want)
    sd = net.state_dict()
    for key in [k for k in golden.files if k.startswith(name + "_after_")]:
        pk = key[len(name) + 7:]
        assert np.allclose(sd[pk].cpu().numpy(), golden[key], rtol=1e-4, atol=1e-5), pk


@pytest.mark.parametrize("name", ["small", "ref53", "bl"])
def test_generate_eval_path(cuda, golden, name):
    net = H.make_product(name, "test", golden).to(cuda)
    net.eval()
    data = H.make_data(name, cuda)
    with torch.no_grad():
        ep, eval_dict, parsed = net.generate(data)
    _check_endpoints(ep, golden, name + "_gen_")
    assert np.array_equal(eval_dict["pred_mask"], golden["%s_gen_pred_mask" % name])
    assert np.abs(parsed["pred_corners_3d"] - golden["%s_gen_corners" % name]).max() < 1e-4
    assert np.abs(parsed["obj_prob"] - golden["%s_gen_obj_prob" % name]).max() < 1e-5
    assert [len(x) for x in eval_dict["batch_pred_map_cls"]] == golden["%s_gen_npred" % name].tolist()
    assert len(eval_dict["batch_gt_map_cls"]) == data["input_joints"].shape[0]
    assert set(ep["pi"]) == {"center", "size", "heading"}


def test_product_vs_oracle_fresh_inputs(cuda, golden):
    """Beyond the committed goldens: a new seed, product (GPU) vs the CPU oracle port, train step."""
    from oracle.model_ref import RefP2RNet
    from pose2room_b200 import synthetic
    name = "small"
    B, T, J, S, P = H.CONFIGS[name]
    net = H.make_product(name, "train", golden).to(cuda)
    net.train()
    ref = RefP2RNet({k: v.cpu() for k, v in net.state_dict().items()}, joint_num=J, num_seeds=S, num_target=P)
    data = synthetic.make_batch(3, 200, J, seed=77)
    ep_r = ref.forward(data)
    loss_r = ref.loss(ep_r, data)
    data_g = {k: (v.to(cuda) if isinstance(v, torch.Tensor) else v) for k, v in data.items()}
    ep = net(data_g)
    loss = net.loss(ep, data_g)
    for k in H.EP_KEYS:
        a, b = ep[k].detach().cpu(), ep_r[k].detach()
        if a.dtype in (torch.int64, torch.int32):
            assert torch.equal(a, b), k
        else:
            tol = max(1e-4, 2e-5 * float(b.abs().max())) if k.endswith("_scores") else 1e-4
            assert ((a - b).abs() <= tol).all(), (k, (a - b).abs().max())
    assert abs(loss["total"].item() - loss_r["total"].item()) < 1e-4 * abs(loss_r["total"].item())


def test_cpu_input_fails_loudly(golden):
    net = H.make_product("small", "train", golden)
    with pytest.raises(RuntimeError, match="no CPU path"):
        net(H.make_data("small"))


def test_bf16_throughput_mode_tracks_fp32_reference(cuda, golden):
    """Throughput mode (bf16 activations, tcgen05 GEMMs, fp32 accumulate / BN statistics / losses) against the
    fp32 reference goldens.  bf16 rounding (2^-9 relative per activation) makes this a statistical check: the seed
    sampling is untouched (exact), the loss must agree to a few percent, every output must be finite, and the box
    tensors must stay within 5e-2 wherever the FPS picks (which depend on vote_xyz) coincide."""
    from pose2room_b200 import gemm_sm100
    gemm_sm100.install()
    try:
        for name in ["small", "bl"]:
            net = H.make_product(name, "train", golden, precision="bf16").to(cuda)
            net.train()
            data = H.make_data(name, cuda)
            ep = net(data)
            assert np.array_equal(ep["seed_inds"].cpu().numpy(), golden[name + "_train_seed_inds"])
            loss = net.loss(ep, data)
            # vote loss depends on the backbone + voting MLP only: tight.  The box regression losses also depend on
            # which votes FPS picked (bf16 noise can swap picks): loose.  Objectness / class CE ride on the fixture's
            # deliberately sharp logits and are not compared.
            want = float(golden["%s_loss_vote_loss" % name])
            assert abs(loss["vote_loss"].item() - want) < 0.03 * abs(want), (name, loss["vote_loss"].item(), want)
            reg = sum(loss[k].item() for k in ["center_loss", "size_loss", "heading_loss"])
            reg_want = sum(float(golden["%s_loss_%s" % (name, k)]) for k in ["center_loss", "size_loss", "heading_loss"])
            assert abs(reg - reg_want) < 0.25 * reg_want, (name, reg, reg_want)
            assert torch.isfinite(loss["total"])
            loss["total"].backward()
            for k, p in net.named_parameters():
                assert p.grad is None or torch.isfinite(p.grad).all(), k
            got_inds = ep["aggregated_vote_inds"].cpu().numpy()
            want_inds = golden[name + "_train_aggregated_vote_inds"]
            # FPS is sequential: one swapped pick changes the tail, so compare the picked SETS per scene
            overlap = np.mean([len(set(g.tolist()) & set(w.tolist())) / float(len(w)) for g, w in zip(got_inds, want_inds)])
            assert overlap > 0.5, overlap
            same = got_inds == want_inds
            if same.all():
                for k in ["center", "size", "heading"]:
                    err = np.abs(ep[k].detach().cpu().numpy() - golden["%s_train_%s" % (name, k)]).max()
                    assert err < 5e-2, (name, k, err)
    finally:
        gemm_sm100.uninstall()


def test_overlapped_weight_gradients_match_plain_backward(cuda, golden):
    """ops.overlap_weight_grads() (dW / db on a side stream, handed to autograd at the join) must give the same
    gradients as a plain loss.backward(): same kernels, different schedule."""
    from pose2room_b200 import gemm_sm100, ops
    gemm_sm100.install()
    try:
        grads = []
        for use_overlap in [False, True]:
            from pose2room_b200 import synthetic
            net = H.make_product("bl", "train", golden, precision="bf16").to(cuda)
            net.train()
            data = {k: (v.to(cuda) if isinstance(v, torch.Tensor) else v)
                    for k, v in synthetic.make_batch(4, 1024, 25, seed=3).items()}   # B=4: every layer is large enough to defer
            if use_overlap:
                with ops.overlap_weight_grads():
                    ep = net(data)
                    loss = net.loss(ep, data)["total"]
                    loss.backward()
            else:
                ep = net(data)
                loss = net.loss(ep, data)["total"]
                loss.backward()
            torch.cuda.synchronize()
            grads.append({k: p.grad.detach().double().clone() for k, p in net.named_parameters() if p.grad is not None})
        assert set(grads[0]) == set(grads[1])
        for k in grads[0]:
            a, b = grads[0][k], grads[1][k]
            scale = a.abs().max().item() + 1e-12
            assert (a - b).abs().max().item() <= 2e-3 * scale, (k, (a - b).abs().max().item(), scale)
    finally:
        gemm_sm100.uninstall()


def test_deferred_gradients_direct_and_through_autograd_agree(cuda, golden):
    """At the join of ops.overlap_weight_grads() a parameter takes its deferred gradient directly (no AccumulateGrad copy);
    with that switched off, or with a tensor hook on the parameter, it goes through autograd.  Same gradients either way,
    the hook fires, and a second backward without zero_grad accumulates."""
    from pose2room_b200 import gemm_sm100, ops, synthetic
    gemm_sm100.install()
    saved = ops._DIRECT_LEAF_GRADS
    try:
        out = []
        for direct in (True, False):
            ops._DIRECT_LEAF_GRADS = direct
            net = H.make_product("bl", "train", golden, precision="bf16").to(cuda)
            net.train()
            data = {k: (v.to(cuda) if isinstance(v, torch.Tensor) else v)
                    for k, v in synthetic.make_batch(4, 1024, 25, seed=5).items()}
            fired = []
            hooked = net.backbone.st_gcn_networks[2].gcn.conv.weight
            hooked.register_hook(lambda g: fired.append(g.shape))
            for _ in range(2):                       # two backward passes, no zero_grad in between: gradients add up
                torch.manual_seed(11)
                with ops.overlap_weight_grads():
                    loss = net.loss(net(data), data)["total"]
                    loss.backward()
            torch.cuda.synchronize()
            assert len(fired) == 2, "the tensor hook of a deferred-gradient parameter did not fire"
            out.append({k: p.grad.detach().double().clone() for k, p in net.named_parameters() if p.grad is not None})
        assert set(out[0]) == set(out[1])
        for k in out[0]:
            a, b = out[0][k], out[1][k]
            scale = a.abs().max().item() + 1e-12
            assert (a - b).abs().max().item() <= 2e-3 * scale, (k, (a - b).abs().max().item(), scale)
    finally:
        ops._DIRECT_LEAF_GRADS = saved
        gemm_sm100.uninstall()


def test_weight_shadows_change_nothing(cuda, golden):
    """ops.register_weight_shadows (one multi-tensor fp32 -> bf16 copy per step instead of a cast per layer): the first
    step's loss bit for bit and its gradients to atomics noise; after an optimiser step the shadows must hold the NEW
    weights: seed features and votes of the second step agree with the layer-by-layer conversion (and differ clearly from
    the first step's: the weights did move).  The total loss of step 2 is not compared -- the fixture's sharp logits turn
    one flipped FPS pick (atomics noise of 1e-7 in a weight is enough) into a jump of the cross-entropy terms."""
    from pose2room_b200 import gemm_sm100, ops
    gemm_sm100.install()
    try:
        out, counters = [], []
        for use_shadows in (False, True):
            net = H.make_product("small", "train", golden, precision="bf16").to(cuda).train()
            if use_shadows:
                ops.register_weight_shadows(net)
            opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], lr=1e-3)
            data = H.make_data("small", cuda)
            rec = []
            for _ in range(2):
                opt.zero_grad(set_to_none=True)
                with ops.overlap_weight_grads():
                    ep = net(data)
                    loss = net.loss(ep, data)["total"]
                    loss.backward()
                rec.append((loss.detach().clone(), {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None},
                            ep["seed_features"].detach().double().clone(), ep["vote_xyz"].detach().double().clone()))
                opt.step()
            out.append(rec)
            ops.clear_weight_shadows()
            counters.append(sorted((k, int(v)) for k, v in net.state_dict().items() if k.endswith("num_batches_tracked")))
        assert counters[0] == counters[1] and all(v == 2 for _, v in counters[0])    # one increment per BatchNorm per step
        (la, ga, sfa, va), (lb, gb, sfb, vb) = out[0][0], out[1][0]
        assert torch.equal(la, lb)                     # the forward pass is deterministic: bit-identical first loss
        assert set(ga) == set(gb)
        for k in ga:    # (weight gradients that end in fp32 atomics -- K = 3 first layers, split-K -- differ in the last bits)
            scale = float(ga[k].abs().max()) + 1e-12
            assert float((ga[k].double() - gb[k].double()).abs().max()) <= 1e-4 * scale, k
        (_, _, sfa2, va2), (_, _, sfb2, vb2) = out[0][1], out[1][1]
        rel = lambda x, y: float((x - y).norm() / y.norm())
        moved = rel(sfa2, sfa)
        assert moved > 1e-3, moved                     # the optimiser step changed the features ...
        assert rel(sfb2, sfa2) <= 0.05 * moved and rel(vb2, va2) <= 0.05 * max(rel(va2, va), 1e-4), (rel(sfb2, sfa2), moved)
    finally:
        ops.clear_weight_shadows()
        gemm_sm100.uninstall()


def _fused_vs_chain_on_random_predictions(dev):
    """Fused loss kernel vs the chain of torch kernels on predictions that hit every branch (near / far / in-between
    proposals, both sides of the huber knee, masked seeds): ten numbers and the six input gradients."""
    import os
    from pose2room_b200.config import P2RConfig
    from pose2room_b200.p2rnet.loss import BoxNetDetectionLoss
    from tests.test_loss_math import make_case
    crit = BoxNetDetectionLoss(1, 0, P2RConfig(mode="train", joint_num=25))
    for seed, hd in [(1, torch.float64), (3, torch.float32)]:
        est, gt, sem_obj = make_case(seed, B=4, T=64, S=40, P=32, heading_dtype=hd)
        results = []
        before = os.environ.get("P2R_FUSED_LOSS")
        for flag in ("0", "1"):
            os.environ["P2R_FUSED_LOSS"] = flag
            leaves = {k: est[k].clone().to(dev).requires_grad_(True) for k in ("vote_xyz", "center", "size", "heading")}
            so = sem_obj.clone().to(dev).requires_grad_(True)
            e = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in est.items()}
            e.update(leaves)
            e["objectness_scores"], e["sem_cls_scores"] = so[..., 0:2], so[..., 2:]
            g = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in gt.items()}
            out = crit(e, g, None)
            out["total"].backward()
            torch.cuda.synchronize()
            results.append(({k: v.item() for k, v in out.items()},
                            dict({k: v.grad.double().cpu() for k, v in leaves.items()}, sem_obj=so.grad.double().cpu()),
                            {k: v.dtype for k, v in out.items()}))
        if before is None:
            os.environ.pop("P2R_FUSED_LOSS")
        else:
            os.environ["P2R_FUSED_LOSS"] = before
        (a, ga, da), (b, gb, db) = results
        assert da == db, (da, db)
        for k in a:
            assert abs(a[k] - b[k]) <= 2e-6 * max(1.0, abs(a[k])), (k, a[k], b[k])
        for k in ga:
            assert (ga[k] - gb[k]).abs().max().item() <= 2e-6 * max(1e-3, ga[k].abs().max().item()), k


def _fused_gmm_vs_torch_path(dev):
    """Fused mixture-head kernels vs the module's torch path on the same eps (real sigma, f32 and f64 heads, f32 and bf16
    logits): prediction and the gradients of logits / mu / log_sigma."""
    import os
    from pose2room_b200.p2rnet.mdn import MixtureDensityHead, Struct, _FusedGMMPredict
    for G, D, mu_dtype, lg_dtype in [(100, 3, torch.float32, torch.float32), (100, 2, torch.float64, torch.float32),
                                     (100, 3, torch.float32, torch.bfloat16), (33, 2, torch.float64, torch.bfloat16)]:
        rows = 4096 + 5
        gen = torch.Generator().manual_seed(G + D)
        head = MixtureDensityHead(Struct(input_dim=8, num_gaussian=G, out_dim=D, n_samples=1, central_tendency="mean",
                                         mu_bias_init=torch.randn(G, D, generator=gen).to(mu_dtype))).to(dev)
        with torch.no_grad():
            head.log_sigma.copy_((0.5 * torch.randn(G, D, generator=gen) - 0.5).to(dev))
        base = (2.0 * torch.randn(rows, G, generator=gen) - 1.0).to(lg_dtype).to(dev)
        dout = torch.randn(rows, D, generator=gen).to(mu_dtype).to(dev)
        res = []
        for fused in (False, True):
            for p in head.parameters():
                p.grad = None
            logits = base.clone().requires_grad_(True)
            torch.manual_seed(11)
            if fused:
                eps = head.mu.data.new(rows, G, 1, D).normal_()
                out = _FusedGMMPredict.apply(logits, head.mu, head.log_sigma, eps)
            else:
                out = head.point_prediction(torch.sigmoid(logits.float()))
            out.backward(dout)
            torch.cuda.synchronize()
            res.append([t.detach().double().cpu() for t in (out, logits.grad, head.mu.grad, head.log_sigma.grad)] + [out.dtype])
        assert res[0][4] == res[1][4]
        lg_tol = 1e-2 if lg_dtype == torch.bfloat16 else 3e-6        # d logits is rounded to bf16 on the fused path
        for name, a, b, tol in zip(("out", "dlogits", "dmu", "dls"), res[0], res[1], (3e-6, lg_tol, 2e-5, 2e-5)):
            assert (a - b).abs().max().item() <= tol * max(1e-3, a.abs().max().item()), (G, D, name, (a - b).abs().max().item())


def _run_isolated(flags, extra=""):
    import os
    import subprocess
    import sys
    code = ("import os, torch, tests.test_model_gpu as T, tests.model_helpers as H; dev = torch.device('cuda:0'); "
            "torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False; " + extra +
            "".join("os.environ['%s'] = '1'; " % f for f in flags) + "g = H.load_golden(); "
            "[T.test_train_forward_loss_backward(dev, g, n) for n in ('small', 'ref53', 'bl')]; "
            "T.test_bf16_throughput_mode_tracks_fp32_reference(dev, g); print('ISOLATED-OK')")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ), capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "ISOLATED-OK" in r.stdout, (r.stdout[-500:], r.stderr[-2000:])


def _fused_vote_vs_torch_path(dev):
    """CenterVoteModule with the fused tail vs its torch path + the normalisation of P2RNet._trunk (fp32 and bf16)."""
    import os
    from pose2room_b200.config import P2RConfig
    from pose2room_b200.p2rnet.vote_center import CenterVoteModule
    for precision in ("fp32", "bf16"):
        if precision == "bf16":
            from pose2room_b200 import gemm_sm100
            gemm_sm100.install()
        try:
            torch.manual_seed(0)
            mod = CenterVoteModule(P2RConfig(mode="train", joint_num=25, precision=precision)).to(dev).train()
            skel = torch.randn(4, 512, 25, 3, device=dev)
            base = torch.randn(4, 512, 256, device=dev)
            g_xyz, g_feat = torch.randn(4, 512, 3, device=dev), torch.randn(4, 512, 256, device=dev)
            res = []
            for fused in (False, True):
                mod.zero_grad(set_to_none=True)
                sf = base.clone().requires_grad_(True)
                if fused:
                    xyz, feat = mod(skel, sf, normalize=True)
                else:
                    xyz, feat = mod(skel, sf)
                    feat = feat.div(torch.norm(feat, p=2, dim=2).unsqueeze(2))
                torch.autograd.backward([xyz, feat], [g_xyz, g_feat])
                torch.cuda.synchronize()
                res.append([t.detach().double().cpu() for t in (xyz, feat, sf.grad, mod.conv_input[2].conv.weight.grad,
                                                                mod.conv_input[0].conv.weight.grad)])
            tol = 2e-2 if precision == "bf16" else 1e-5      # bf16: d net is rounded to bf16 on the fused path
            for name, a, b in zip(("xyz", "feat", "d_seed_features", "dW2", "dW0"), res[0], res[1]):
                assert (a - b).abs().max().item() <= tol * max(1e-3, a.abs().max().item()), (precision, name)
        finally:
            if precision == "bf16":
                gemm_sm100.uninstall()


def test_fused_loss_kernel_vs_torch_chain(cuda):
    _fused_vs_chain_on_random_predictions(cuda)


def test_fused_mixture_head_kernels_vs_torch_path(cuda):
    _fused_gmm_vs_torch_path(cuda)


def test_fused_vote_tail_kernels_vs_torch_path(cuda):
    _fused_vote_vs_torch_path(cuda)


def test_unfused_paths_pass_the_parity_tests(cuda):
    """The fused detection-loss / mixture-head / vote-tail kernels are the default since round 2, so every parity test of
    this file runs them.  Here, in a process of its own, the reference-golden parity tests and the bf16 check with all
    three flags OFF: the chains of torch kernels the fused kernels replaced stay a tested second implementation."""
    _run_isolated([], "os.environ.update(P2R_FUSED_LOSS='0', P2R_FUSED_GMM='0', P2R_FUSED_VOTE='0'); ")
