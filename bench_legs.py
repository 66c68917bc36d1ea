"""Secondary legs of bench.py (one GPU, each in a child process of the supervisor, after the headline line is safe).

  python bench.py --leg gpu_reference   "the reference on B200" (second baseline): the reference's own
                                        computation -- fp32 channel-first Conv1d/Conv2d + einsum through cuDNN / cuBLAS,
                                        materialised grouping, per-sample loss loop (oracle/model_ref.py, the port that is
                                        pinned against the unmodified reference by tests/golden/p2rnet.npz) -- on the GPU
                                        with the UNMODIFIED reference pointnet2 kernels compiled for sm_100a
                                        (oracle/_ref/p2r_ref_ext.so).  Train step and forward only, TF32 off and on.
  python bench.py --leg extras          BASELINE.json config #2 (forward only at B=32: whole model + the set-abstraction
                                        operators one by one, ours vs the reference kernels) and config #5 (eval: 1000
                                        scenes through decode / far-box / NMS / AP, ms per scene, mAP@0.25 / 0.5 next to
                                        the reference's on the same inputs; the reference's host path timed on a sample).

Nothing here is the product: oracle/ is used as the measured BASELINE (allowed for bench.py's baseline legs), never as
the thing shipped.  /root/reference is not read (it does not exist on the GPU box)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
B_PER_GPU, T_FRAMES, JOINTS = 32, 1024, 25


def _events_ms(fn, iters, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def _median_ms(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def _state_dict():
    from pose2room_b200 import synthetic
    from pose2room_b200.config import P2RConfig
    from pose2room_b200.p2rnet import P2RNet
    torch.manual_seed(42)
    np.random.seed(42)
    template = P2RNet(P2RConfig(mode="train", joint_num=JOINTS, num_frames=T_FRAMES)).state_dict()
    return synthetic.deterministic_state_dict(template, seed=7)


# ------------------------------------------------------------------------------------------ the reference on B200
def gpu_reference(batch=B_PER_GPU, steps=5, warmup=2):
    from oracle import build_ref_ext
    from oracle.model_ref import RefP2RNet
    from pose2room_b200 import synthetic
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ext = build_ref_ext.load_ref_ext()
    if ext is None:
        return {"unavailable": "oracle/_ref/p2r_ref_ext.so is not in the snapshot (it is built where /root/reference exists)"}
    sd = _state_dict()
    data = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v)
            for k, v in synthetic.make_batch(batch, T_FRAMES, JOINTS, seed=1234).items()}
    out = {"what": "reference formulation (fp32 conv + einsum via cuDNN/cuBLAS, oracle/model_ref.py) + UNMODIFIED reference "
                   "pointnet2 kernels for sm_100a (oracle/_ref/p2r_ref_ext.so), eager like the reference's trainer; "
                   "train step = forward + loss + backward + AdamW, B=%d, T=%d, J=%d" % (batch, T_FRAMES, JOINTS),
           "batch": batch, "steps": steps, "warmup": warmup, "unit": "sequences/s"}
    for tag, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        net = RefP2RNet(sd, joint_num=JOINTS, num_seeds=512, num_target=128, training=True, device=dev, ext=ext)
        opt = torch.optim.AdamW(net.parameters(), lr=1e-3)

        def step():
            opt.zero_grad(set_to_none=True)
            ep = net.forward(data)
            loss = net.loss(ep, data)["total"]
            loss.backward()
            opt.step()
            return loss

        def forward_only():
            with torch.no_grad():
                return net.forward(data)
        first = float(step())
        torch.cuda.reset_peak_memory_stats()
        ms = _events_ms(step, steps, warmup)
        ms_f = _events_ms(forward_only, steps, warmup)
        out[tag] = {"train_ms_per_step": ms, "train_sequences_per_s": batch / (ms * 1e-3), "forward_ms": ms_f,
                    "forward_sequences_per_s": batch / (ms_f * 1e-3), "first_step_loss": first,
                    "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}
        del net, opt
        torch.cuda.empty_cache()
    out["note"] = ("tf32: torch.backends.{cuda.matmul,cudnn}.allow_tf32 = True, the defaults of the reference's pinned "
                   "torch 1.8.1 (environment.yml) on an Ampere-or-later GPU; fp32: both off (the setting of the 1e-4 parity tests)")
    return out


# ------------------------------------------------------------------------------------------ config #2: forward only
def _product(precision, dev, mode="train"):
    from pose2room_b200.config import P2RConfig
    from pose2room_b200.p2rnet import P2RNet
    torch.manual_seed(42)
    np.random.seed(42)
    net = P2RNet(P2RConfig(mode=mode, joint_num=JOINTS, num_frames=T_FRAMES, precision=precision))
    net.load_state_dict(_state_dict())
    return net.to(dev)


def forward_only(batch=B_PER_GPU, iters=10):
    """BASELINE.json config #2: synthetic B=32, T=1024, J=25, forward only (train-mode forward = batch-statistics
    BatchNorm and sampled mixture heads, no autograd graph), whole model, bf16 and fp32 modes, as a CUDA graph."""
    from pose2room_b200 import gemm_sm100, synthetic
    dev = torch.device("cuda", 0)
    data = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v)
            for k, v in synthetic.make_batch(batch, T_FRAMES, JOINTS, seed=1234).items()}
    out = {"what": "P2RNet forward only (train-mode forward under no_grad), B=%d, T=%d, J=%d, one CUDA graph" % (batch, T_FRAMES, JOINTS)}
    for precision in ("bf16", "fp32"):
        if precision == "bf16":
            gemm_sm100.install()
        try:
            net = _product(precision, dev).train()

            def fwd():
                with torch.no_grad():
                    return net(data)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    fwd()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graphed = True
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    fwd()
                run = g.replay
            except Exception as e:      # say so, do not hide it
                graphed, run = False, fwd
                print("bench_legs: forward-only capture failed (%r), timing eagerly" % (e,), file=sys.stderr)
            ms = _events_ms(run, iters, 3)
            out[precision] = {"forward_ms": ms, "forward_sequences_per_s": batch / (ms * 1e-3), "cuda_graph": graphed}
            del net
        finally:
            if precision == "bf16":
                gemm_sm100.uninstall()
    return out


def sa_operators():
    """The set-abstraction operators of config #2 one by one (CUDA events, median of 20), ours vs the UNMODIFIED reference
    kernels: the live shape of ProposalNet (32 x 512 votes -> 128 proposals, r 0.3, 16 samples, 256 channels) and the
    VoteNet-sized cloud of SURVEY 8d (32 x 25600 points -> 2048, r 0.2, 64 samples, 64 channels; not a reference config)."""
    from oracle import build_ref_ext
    from pose2room_b200 import ext, synthetic
    dev = torch.device("cuda", 0)
    ref = build_ref_ext.load_ref_ext()
    rows = []
    for name, (B, N, M, r, ns, C) in {"live": (32, 512, 128, 0.3, 16, 256), "large": (32, 25600, 2048, 0.2, 64, 64)}.items():
        xyz = torch.from_numpy(synthetic.make_cloud(B, N, seed=1)).to(dev)
        feats = torch.randn(B, C, N, device=dev)
        impls = {"ours": ext}
        if ref is not None:
            impls["reference"] = ref
        for iname, mod in impls.items():
            idx = mod.furthest_point_sampling(xyz, M)
            new_xyz = mod.gather_points(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
            bq = mod.ball_query(new_xyz, xyz, r, ns)
            go = torch.randn(B, C, M, ns, device=dev)
            few = 5 if N > 1000 else 20
            rows.append(dict(shape=name, impl=iname,
                             fps_us=1e3 * _median_ms(lambda: mod.furthest_point_sampling(xyz, M), iters=few),
                             ball_query_us=1e3 * _median_ms(lambda: mod.ball_query(new_xyz, xyz, r, ns), iters=few),
                             group_us=1e3 * _median_ms(lambda: mod.group_points(feats, bq), iters=few),
                             group_grad_us=1e3 * _median_ms(lambda: mod.group_points_grad(go, bq, N), iters=few),
                             three_nn_us=1e3 * _median_ms(lambda: mod.three_nn(xyz, new_xyz), iters=few)))
    return rows


def sa_module(batch=B_PER_GPU):
    """The whole set-abstraction layer of ProposalNet (FPS -> gather -> ball query -> group -> 256-256-256 shared MLP -> max
    over 16) forward at the live shape: the product's path (ProposalNet._aggregate, bf16 and fp32) vs the reference's
    (reference kernels + cuDNN 1x1 convs + max_pool2d over a materialised (B,256,128,16) tensor, fp32)."""
    from oracle import build_ref_ext
    from oracle.model_ref import RefP2RNet
    from pose2room_b200 import gemm_sm100, synthetic
    dev = torch.device("cuda", 0)
    xyz = torch.from_numpy(synthetic.make_cloud(batch, 512, seed=5)).to(dev)
    feats = torch.randn(batch, 512, 256, device=dev)
    feats = feats / feats.norm(dim=2, keepdim=True)
    out = {}
    for precision in ("bf16", "fp32"):
        if precision == "bf16":
            gemm_sm100.install()
        try:
            det = _product(precision, dev).detection

            def run():
                with torch.no_grad():
                    return det._aggregate(xyz, feats)
            out["ours_" + precision + "_us"] = 1e3 * _median_ms(run)
        finally:
            if precision == "bf16":
                gemm_sm100.uninstall()
    ext = build_ref_ext.load_ref_ext()
    if ext is not None:
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        ref = RefP2RNet(_state_dict(), joint_num=JOINTS, training=True, device=dev, ext=ext)
        fcf = feats.transpose(1, 2).contiguous()

        def run_ref():
            with torch.no_grad():
                return ref._sa(xyz, fcf)
        out["reference_fp32_us"] = 1e3 * _median_ms(run_ref)
    return out


# ------------------------------------------------------------------------------------------ config #5: eval
def eval_1k(chunk=125, ref_sample=16):
    """BASELINE.json config #5: the 1000-scene evaluation set of tests/golden/eval1k.npz (inputs rebuilt from (seed, index),
    outputs of the UNMODIFIED reference recorded in the fixture) through the product's eval path: decode + far-box + 3-D NMS
    kernels, assembly, then APCalculator (OBB-IoU kernel + host matching) at IoU 0.25 and 0.5.  Predictions start on the
    device (they are network outputs); the timed region ends with numpy results on the host, like the reference's."""
    from oracle import geometry_ref as G
    from pose2room_b200 import ap_helper, synthetic
    from pose2room_b200.config import P2RConfig
    dev = torch.device("cuda", 0)
    cfg = P2RConfig(mode="test", joint_num=JOINTS).eval_config
    g = np.load(os.path.join(ROOT, "tests", "golden", "eval1k.npz"))
    n, seed = int(g["n_scenes"]), int(g["seed"])
    batches = []
    for start in range(0, n, chunk):
        est, gt = synthetic.make_eval_batch(seed, start, min(chunk, n - start))
        batches.append(({k: v.to(dev) for k, v in est.items()}, gt, {"input_joints": gt["input_joints"].to(dev)}))
    ap_helper.parse_predictions(batches[0][0], batches[0][2], cfg)      # warm-up (module load, allocator)
    torch.cuda.synchronize()
    preds, gts, t_parse = [], [], 0.0
    for est, gt, data in batches:
        t0 = time.perf_counter()
        eval_dict, parsed = ap_helper.parse_predictions(est, data, cfg)
        eval_dict = ap_helper.assembly_pred_map_cls(eval_dict, parsed, cfg)
        gt_map = ap_helper.assembly_gt_map_cls(ap_helper.parse_groundtruths(gt, cfg))
        t_parse += time.perf_counter() - t0
        preds += eval_dict["batch_pred_map_cls"]
        gts += gt_map
    metrics, t0 = {}, time.perf_counter()
    for thr in (0.25, 0.5):
        calc = ap_helper.APCalculator(thr)
        calc.step(preds, gts)
        metrics[thr] = calc.compute_metrics()
    t_ap = time.perf_counter() - t0
    # the reference's host path (numpy decode loops + scipy Delaunay far-box test + numpy NMS: oracle/geometry_ref.py, which
    # reproduces the unmodified reference bit for bit on this set) on a sample of the same scenes, here, now
    est, gt = synthetic.make_eval_batch(seed, 0, ref_sample)
    hip = gt["input_joints"][:, :, 0].numpy()
    t0 = time.perf_counter()
    G.parse_predictions(est["center"].numpy(), est["size"].numpy(), est["heading"].numpy(), est["objectness_scores"].numpy(),
                        est["sem_cls_scores"].numpy(), hip)
    t_ref = (time.perf_counter() - t0) / ref_sample
    return {"what": "1000 scenes x 128 proposals: parse_predictions (decode / far-box / 3-D NMS kernels) + assembly, then AP at "
                    "IoU 0.25 and 0.5 (OBB-IoU kernel + host matching)",
            "scenes": n, "parse_ms_per_scene": 1e3 * t_parse / n, "ap_ms_per_scene": 1e3 * t_ap / n,
            "total_ms_per_scene": 1e3 * (t_parse + t_ap) / n,
            "mAP@0.25": metrics[0.25]["mAP"], "mAP@0.5": metrics[0.5]["mAP"],
            "reference_mAP@0.25": float(g["map_25"]), "reference_mAP@0.5": float(g["map_50"]),
            "cpu_reference_parse_ms_per_scene": 1e3 * t_ref,
            "cpu_reference_sample": "%d scenes, parse_predictions only, %d host threads available" % (ref_sample, os.cpu_count() or 1),
            "recorded_reference": {"parse_ms_per_scene": 1e3 * float(g["parse_seconds"]) / n,
                                   "ap_ms_per_scene": 1e3 * (float(g["ap_seconds_25"]) + float(g["ap_seconds_50"])) / n,
                                   "where": "the UNMODIFIED reference in the build container (8 host threads, Pool(10)), "
                                            "tests/golden/make_golden_eval1k.py"}}


def extras():
    from pose2room_b200 import _lib
    torch.cuda.set_device(0)
    _lib.load()
    out = {}
    for name, fn in (("forward_only", forward_only), ("sa_module_forward", sa_module), ("sa_operators", sa_operators),
                     ("eval_1k", eval_1k)):
        try:
            out[name] = fn()
        except Exception as e:      # informational legs: report, never lose the others
            out[name] = {"error": repr(e)}
    return out


def main(leg):
    res = {"gpu_reference": gpu_reference, "extras": extras}[leg]()
    print(json.dumps(res), flush=True)
