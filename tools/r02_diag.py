"""Round-2 first GPU call: name the assertion behind each of the two round-1 B200 failures.

  A. the reference's demo sequence through the product (tests/test_demo_sequence.py) -- per-key differences against the
     reference goldens and, stage by stage, against the CPU oracle run on the same box;
  B. the fused loss / mixture-head / vote-tail kernels: each kernel against the path it replaces, then the golden parity
     tests with ONE flag at a time.
Everything prints; nothing asserts.  Output: gpurun_out/r02_diag.log (stdout)."""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")


def section(name):
    print("\n" + "=" * 20 + " " + name + " " + "=" * 20, flush=True)


def diff(name, got, want):
    got, want = np.asarray(got), np.asarray(want)
    if got.shape != want.shape:
        print("  %-24s SHAPE %s vs %s" % (name, got.shape, want.shape))
        return
    if want.dtype.kind in "iub":
        bad = got != want
        print("  %-24s %s  mismatches %d / %d%s" % (name, "EXACT" if not bad.any() else "DIFF ", int(bad.sum()), bad.size,
                                                   "" if not bad.any() else "  first at %s got %s want %s" % (
                                                       np.argwhere(bad)[0].tolist(), got[bad][0], want[bad][0])))
    else:
        d = np.abs(got.astype(np.float64) - want.astype(np.float64))
        print("  %-24s max|d| %.3e  at %s  (|want| max %.3e, dtype %s/%s, nan %d)" % (
            name, float(np.nanmax(d)) if d.size else 0.0, np.unravel_index(int(np.nanargmax(d)), d.shape) if d.size else (),
            float(np.abs(want).max()) if want.size else 0.0, got.dtype, want.dtype, int(np.isnan(got).sum())))


def demo():
    section("A. demo sequence")
    from tests import test_demo_sequence as D
    from oracle.model_ref import RefP2RNet
    g = np.load(D.GOLDEN)
    net = D._product("test")
    sd = D._weights(g, net.state_dict())
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    data = D._inputs(g)
    gdata = dict(data, input_joints=data["input_joints"].to(dev))
    with torch.no_grad():
        ep, eval_dict, parsed = net.generate(gdata, eval=False)
    print("product vs the unmodified reference's goldens:")
    for k in D.EP_KEYS:
        diff(k, ep[k].detach().cpu().numpy(), g["gen_" + k])
    diff("pred_mask", eval_dict["pred_mask"], g["gen_pred_mask"])
    diff("corners", parsed["pred_corners_3d"], g["gen_corners"])
    print("  npred", [len(x) for x in eval_dict["batch_pred_map_cls"]], g["gen_npred"].tolist())
    # stage by stage against the oracle on this box's CPU
    ref = RefP2RNet(sd, joint_num=D.J, num_seeds=D.S, num_target=D.P, training=False)
    ep_r, parsed_r = ref.generate(data)
    print("product vs the CPU oracle (same box):")
    for k in sorted(ep_r):
        if k in ep and isinstance(ep_r[k], torch.Tensor):
            diff(k, ep[k].detach().cpu().numpy(), ep_r[k].detach().numpy())
    # teacher-forced detection head: the oracle's votes into the product's aggregation
    with torch.no_grad():
        vx, vf = ep_r["vote_xyz"].to(dev), ep_r["vote_features"].to(dev)
        ep2, _ = net.detection.generate(vx, vf, {}, False)
    print("product detection head on the ORACLE's votes vs the oracle:")
    for k in ["aggregated_vote_inds", "aggregated_vote_xyz", "center", "size", "heading", "objectness_scores", "sem_cls_scores"]:
        diff(k, ep2[k].detach().cpu().numpy(), ep_r[k].detach().numpy())
    # native ops alone on the oracle's votes
    from pose2room_b200 import ext
    from oracle.pointnet2_ref import RefExt
    xyz = ep_r["vote_xyz"].contiguous()
    a = ext.furthest_point_sampling(xyz.to(dev), D.P).cpu()
    b = RefExt.furthest_point_sampling(xyz, D.P)
    diff("fps(oracle votes)", a.numpy(), b.numpy())
    new_xyz = torch.gather(xyz, 1, b.long()[:, :, None].expand(-1, -1, 3)).contiguous()
    diff("ball_query(oracle votes)", ext.ball_query(new_xyz.to(dev), xyz.to(dev), 0.3, 16).cpu().numpy(),
         RefExt.ball_query(new_xyz, xyz, 0.3, 16).numpy())
    # the product's own votes through both FPS implementations (is the pick difference a vote_xyz rounding effect?)
    xyz_p = ep["vote_xyz"].detach().cpu().contiguous()
    diff("fps(product votes) gpu/cpu", ext.furthest_point_sampling(xyz_p.to(dev), D.P).cpu().numpy(),
         RefExt.furthest_point_sampling(xyz_p, D.P).numpy())
    d = (xyz_p - xyz).abs()
    print("  vote_xyz product-oracle: max %.3e; distinct rows oracle %d product %d" % (
        d.max().item(), len(np.unique(xyz[0].numpy(), axis=0)), len(np.unique(xyz_p[0].numpy(), axis=0))))


def fused():
    section("B. fused kernels vs the paths they replace")
    import tests.test_model_gpu as T
    for name, fn in [("loss", T._fused_vs_chain_on_random_predictions), ("gmm", T._fused_gmm_vs_torch_path),
                     ("vote", T._fused_vote_vs_torch_path)]:
        for k in ("P2R_FUSED_LOSS", "P2R_FUSED_GMM", "P2R_FUSED_VOTE"):
            os.environ[k] = "0"
        try:
            fn(dev)
            print("  fused %-5s vs replaced path: OK" % name, flush=True)
        except Exception:
            print("  fused %-5s vs replaced path: FAILED" % name)
            traceback.print_exc(file=sys.stdout)
            sys.stdout.flush()
    import tests.model_helpers as H
    g = H.load_golden()
    for flag in ("P2R_FUSED_LOSS", "P2R_FUSED_GMM", "P2R_FUSED_VOTE"):
        for k in ("P2R_FUSED_LOSS", "P2R_FUSED_GMM", "P2R_FUSED_VOTE"):
            os.environ[k] = "1" if k == flag else "0"
        for what in ("small", "ref53", "bl", "bf16"):
            try:
                if what == "bf16":
                    T.test_bf16_throughput_mode_tracks_fp32_reference(dev, g)
                else:
                    T.test_train_forward_loss_backward(dev, g, what)
                print("  %s golden parity %-6s OK" % (flag, what), flush=True)
            except Exception:
                print("  %s golden parity %-6s FAILED" % (flag, what))
                traceback.print_exc(file=sys.stdout)
                sys.stdout.flush()
    for k in ("P2R_FUSED_LOSS", "P2R_FUSED_GMM", "P2R_FUSED_VOTE"):
        os.environ[k] = "0"


if __name__ == "__main__":
    which = sys.argv[1:] or ["demo", "fused"]
    from pose2room_b200 import _lib
    _lib.load()
    for w in which:
        try:
            {"demo": demo, "fused": fused}[w]()
        except Exception:
            traceback.print_exc(file=sys.stdout)
