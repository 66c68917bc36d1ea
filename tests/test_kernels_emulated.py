"""CPU: the fused detection-loss and mixture-head KERNELS and their C-ABI launchers (csrc/loss_ops.cu, csrc/gmm_ops.cu,
unmodified) executed on the host by a minimal CUDA execution-model emulator (tests/csrc/cuda_emu.h: blocks in shuffled
order, threads of a block as pthreads, __syncthreads / warp shuffles / atomics / __threadfence emulated).  What this adds
over test_loss_math.py / test_gmm_math.py (which hold the ARITHMETIC headers to the oracle): the kernels' plumbing -- block
partial sums, the last-block-finalises pattern whatever block finishes last, the block-wide arg-min merge, strided inputs,
tail blocks, the launchers' grid and workspace arithmetic -- against the sequential host harness over the same headers.
Written because these kernels could not be run on a GPU in the session that produced them; not a substitute for it."""
import ctypes
import os.path as osp
import subprocess

import numpy as np
import pytest
import torch

from pose2room_b200 import _lib
from tests import test_gmm_math as TG
from tests import test_loss_math as TL

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "kernels_emu.so")
    inc = [osp.join(ROOT, "include"), osp.join(ROOT, "pose2room_b200", "csrc"), osp.join(ROOT, "tests", "csrc")]
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-pthread", "-shared", "-fPIC", "-std=c++17", "-w",
                    "-DP2R_HOST_EMULATION"] + sum((["-I", i] for i in inc), []) +
                   [osp.join(ROOT, "tests", "csrc", "kernels_emu.cpp"), "-o", so], check=True)
    lib = ctypes.CDLL(so)
    for name in ("p2r_detection_loss", "p2r_detection_loss_grad", "p2r_gmm_mix", "p2r_gmm_mix_grad",
                 "p2r_detection_loss_workspace", "p2r_gmm_mix_workspace", "p2r_vote_tail", "p2r_vote_tail_grad"):
        fn = getattr(lib, name)
        fn.argtypes = _lib.SIGNATURES[name]            # the product's own ctypes signatures
        fn.restype = _lib._RESTYPES.get(name, ctypes.c_int)
    lib.emu_last_error.restype = ctypes.c_char_p
    return lib


@pytest.fixture(scope="module")
def host_loss_lib(tmp_path_factory):
    return TL.host_lib.__wrapped__(tmp_path_factory)


@pytest.fixture(scope="module")
def host_gmm_lib(tmp_path_factory):
    return TG.host_lib.__wrapped__(tmp_path_factory)


def _p(a):
    return a.ctypes.data


def emulated_loss(lib, est, gt, sem_obj, g_total=1.0):
    B, S, J = est["seed_skeleton"].shape[:3]
    T, P, G, C = gt["vote_label"].shape[1], est["center"].shape[1], gt["center_label"].shape[1], sem_obj.shape[2] - 2
    h64 = est["heading"].dtype == torch.float64
    arr = lambda t, dt: np.ascontiguousarray(t.detach().numpy().astype(dt, copy=False))
    so = arr(sem_obj, np.float32)
    a = dict(vote=arr(est["vote_xyz"], np.float32), center=arr(est["center"], np.float32), size=arr(est["size"], np.float32),
             head=arr(est["heading"], np.float64 if h64 else np.float32), agg=arr(est["aggregated_vote_xyz"], np.float32),
             skel=arr(est["seed_skeleton"], np.float32), inds=arr(est["seed_inds"], np.int64),
             vl=arr(gt["vote_label"], np.float32), vm=arr(gt["vote_label_mask"], np.int64),
             gc=arr(gt["center_label"], np.float32), gm=arr(gt["box_label_mask"], np.float32), gs=arr(gt["size"], np.float32),
             gh=arr(gt["heading"], np.float32), gcls=arr(gt["sem_cls_label"], np.int64))
    out32, out64, scales = np.full(8, np.nan, np.float32), np.full(2, np.nan), np.full(4, np.nan)
    hd = np.float64 if h64 else np.float32
    u = [np.full((B, S, 3), np.nan, np.float32), np.full((B, P, 3), np.nan, np.float32), np.full((B, P, 3), np.nan, np.float32),
         np.full((B, P, 3), np.nan, np.float32), np.full((B, P, 2), np.nan, hd), np.full((B, P, 2), np.nan, np.float32),
         np.full((B, P, C), np.nan, np.float32)]
    n_ws = lib.p2r_detection_loss_workspace(B, S)
    ws = np.zeros(n_ws + 4)
    ws[n_ws:] = 777.0                                                   # canary behind the workspace
    rc = lib.p2r_detection_loss(_p(a["vote"]), _p(a["center"]), _p(a["size"]), _p(a["head"]), int(h64), _p(so), 2 + C,
                                _p(so) + 8, 2 + C, _p(a["agg"]), _p(a["skel"]), _p(a["inds"]), _p(a["vl"]), _p(a["vm"]),
                                _p(a["gc"]), _p(a["gm"]), _p(a["gs"]), _p(a["gh"]), _p(a["gcls"]), B, S, J, T, P, G, C, 0,
                                _p(out32), _p(out64), _p(scales), *[_p(x) for x in u], _p(ws), n_ws, None)
    assert rc == 0, lib.emu_last_error()
    assert (ws[n_ws:] == 777.0).all() and int(ws[:1].view(np.uint32)[0]) == B + -(-B * S // 128)   # every block counted once
    g32, g64 = np.zeros(8, np.float32), np.array([0.0, g_total])
    d = [np.full_like(x, np.nan) for x in (u[0], u[1], u[3], u[4], u[5], u[6])]
    rc = lib.p2r_detection_loss_grad(_p(g32), _p(g64), _p(scales), _p(u[0]), _p(u[1]), _p(u[2]), _p(u[3]), _p(u[4]), int(h64),
                                     _p(u[5]), _p(u[6]), B, S, P, C, *[_p(x) for x in d], None)
    assert rc == 0, lib.emu_last_error()
    names = ["vote_xyz", "center", "size", "heading", "objectness_scores", "sem_cls_scores"]
    return out32, out64, dict(zip(names, d))


@pytest.mark.parametrize("shape", [dict(B=3, T=48, S=20, P=16), dict(B=2, T=300, S=200, P=150),
                                   dict(B=8, T=1024, S=512, P=128)])       # the BASELINE shape per sample; 8 of its 32 samples keep the suite short
def test_detection_loss_kernels_under_emulation_equal_the_sequential_harness(emu, host_loss_lib, shape):
    hd = torch.float32 if shape["P"] == 150 else torch.float64
    est, gt, sem_obj = TL.make_case(11, J=25, heading_dtype=hd, **shape)
    want32, want64, want_g = TL.host_loss(host_loss_lib, est, gt, sem_obj)
    out32, out64, grads = emulated_loss(emu, est, gt, sem_obj)
    # same arithmetic header on both sides; only the association of the float64 partial sums differs
    assert np.allclose(out32, want32, rtol=1e-6, atol=0) and np.allclose(out64, want64, rtol=1e-12, atol=0)
    for k in want_g:
        assert not np.isnan(grads[k]).any(), k                           # every element written
        scale = max(1e-30, float(np.abs(want_g[k]).max()))
        assert float(np.abs(grads[k].astype(np.float64) - want_g[k]).max()) <= 1e-6 * scale, k


def test_bad_arguments_are_reported_by_the_launchers(emu):
    assert emu.p2r_detection_loss(*([None] * 4), 1, None, 2, None, 22, *([None] * 10), 2, 8, 25, 16, 4, 65, 22, 0,
                                  *([None] * 11), 0, None) == -1         # G = 65 > P2RL_MAX_GT
    assert b"p2r_detection_loss" in emu.emu_last_error()
    assert emu.p2r_gmm_mix(None, 0, None, 0, None, None, 4, 300, 3, None, None) == -1      # G = 300 > P2RG_MAX_G
    z = np.zeros(8)
    assert emu.p2r_gmm_mix_grad(None, 0, None, 0, None, None, None, 64, 100, 3, None, None, None, _p(z), 8, None) == -1
    assert b"workspace" in emu.emu_last_error()


@pytest.mark.parametrize("G,D,mu_dtype,bf16", [(100, 3, np.float32, False), (100, 2, np.float64, False),
                                               (33, 3, np.float32, True), (256, 4, np.float64, False)])
def test_mixture_kernels_under_emulation_equal_the_sequential_harness(emu, host_gmm_lib, G, D, mu_dtype, bf16):
    rows = 1029 if G == 100 else 77                                      # tail block / tail warps
    rng = np.random.default_rng(G + D)
    logits = (2.0 * rng.normal(size=(rows, G)) - 1.0).astype(np.float32)
    if bf16:                                                             # logits stored as bf16: same values on both sides
        bits = (logits.view(np.uint32) >> 16).astype(np.uint16)
        logits = (bits.astype(np.uint32) << 16).view(np.float32)
    mu = rng.normal(size=(G, D)).astype(mu_dtype)
    ls = (0.5 * rng.normal(size=(G, D)) - 0.5).astype(np.float32)
    eps = rng.normal(size=(rows, G, 1, D)).astype(mu_dtype)
    dout = rng.normal(size=(rows, D)).astype(mu_dtype)
    # sequential harness (float64 storage)
    mu64, eps64, do64 = mu.astype(np.float64), eps.astype(np.float64), dout.astype(np.float64)
    want = np.zeros((rows, D))
    host_gmm_lib.host_gmm_mix(_vp(logits), _vp(mu64), _vp(ls), _vp(eps64), ctypes.c_longlong(rows), G, D, _vp(want))
    w_dlog, w_dmu, w_dls = np.zeros((rows, G), np.float32), np.zeros((G, D)), np.zeros((G, D), np.float32)
    host_gmm_lib.host_gmm_mix_grad(_vp(logits), _vp(mu64), _vp(ls), _vp(eps64), _vp(do64), ctypes.c_longlong(rows), G, D,
                                   _vp(w_dlog), _vp(w_dmu), _vp(w_dls))
    # emulated kernels
    lg_store = bits if bf16 else logits
    out = np.full((rows, D), np.nan, mu_dtype)
    f64 = int(mu_dtype == np.float64)
    assert emu.p2r_gmm_mix(_p(lg_store), int(bf16), _p(mu), f64, _p(ls), _p(eps), rows, G, D, _p(out), None) == 0
    tol = 1e-12 if f64 else 2e-6
    assert np.abs(out.astype(np.float64) - want).max() <= tol * max(1.0, np.abs(want).max())
    n_ws = emu.p2r_gmm_mix_workspace(rows, G, D)
    ws = np.zeros(n_ws + 4)
    ws[n_ws:] = 777.0
    dlog = np.full((rows, G), 0x7fc0 if bf16 else np.nan, np.uint16 if bf16 else np.float32)
    dmu, dls = np.full((G, D), np.nan, mu_dtype), np.full((G, D), np.nan, np.float32)
    assert emu.p2r_gmm_mix_grad(_p(lg_store), int(bf16), _p(mu), f64, _p(ls), _p(eps), _p(dout), rows, G, D, _p(dlog), _p(dmu),
                                _p(dls), _p(ws), n_ws, None) == 0, emu.emu_last_error()
    assert (ws[n_ws:] == 777.0).all() and int(ws[:1].view(np.uint32)[0]) == -(-rows // 32)
    got_dlog = (dlog.astype(np.uint32) << 16).view(np.float32) if bf16 else dlog
    assert np.abs(got_dlog - w_dlog).max() <= (8e-3 if bf16 else 1e-6) * np.abs(w_dlog).max()
    assert np.abs(dmu.astype(np.float64) - w_dmu).max() <= (1e-12 if f64 else 2e-6) * np.abs(w_dmu).max()
    assert np.abs(dls - w_dls).max() <= 2e-6 * np.abs(w_dls).max()


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("bf16", [False, True])
def test_vote_tail_kernels_under_emulation_match_torch_autograd(emu, bf16):
    """vote_ops.cu (kernels + launchers) against the torch expressions they replace (vote_center.py:52-58, network.py:89-90)
    and torch autograd of them; seed_xyz read in place out of a (B,S,J,3) skeleton tensor."""
    B, S, J, C = 3, 43, 5, 256                                           # 129 rows: tail warps in the last block
    gen = torch.Generator().manual_seed(3)
    skel = torch.randn(B, S, J, 3, generator=gen)
    seed_xyz = skel[:, :, 0]                                             # strided view, rows J*3 floats apart
    sf = torch.randn(B, S, C, generator=gen).requires_grad_(True)
    net = torch.randn(B * S, 3 + C, generator=gen)
    if bf16:
        net = net.bfloat16().float()
    net.requires_grad_(True)
    n3 = net.reshape(B, S, 3 + C)
    want_xyz = seed_xyz + n3[..., 0:3]
    v = sf + n3[..., 3:]
    want_feat = v.div(torch.norm(v, p=2, dim=2).unsqueeze(2))
    g_xyz, g_feat = torch.randn(B, S, 3, generator=gen), torch.randn(B, S, C, generator=gen)
    torch.autograd.backward([want_xyz, want_feat], [g_xyz, g_feat])

    rows = B * S
    net_np = net.detach().numpy()
    store = (net_np.view(np.uint32) >> 16).astype(np.uint16) if bf16 else net_np
    skel_np, sf_np = skel.numpy(), sf.detach().numpy()
    xyz, feat, norm = np.full((rows, 3), np.nan, np.float32), np.full((rows, C), np.nan, np.float32), np.full(rows, np.nan, np.float32)
    assert emu.p2r_vote_tail(_p(store), int(bf16), _p(skel_np), J * 3, _p(sf_np), rows, C, _p(xyz), _p(feat), _p(norm), None) == 0
    assert np.abs(xyz - want_xyz.detach().numpy().reshape(rows, 3)).max() <= 1e-6
    assert np.abs(feat - want_feat.detach().numpy().reshape(rows, C)).max() <= 2e-7
    d_net = np.full((rows, 3 + C), 0x7fc0 if bf16 else np.nan, np.uint16 if bf16 else np.float32)
    d_sf = np.full((rows, C), np.nan, np.float32)
    gx, gf = g_xyz.numpy().reshape(rows, 3).copy(), g_feat.numpy().reshape(rows, C).copy()
    assert emu.p2r_vote_tail_grad(_p(gx), _p(gf), _p(feat), _p(norm), rows, C, _p(d_net), int(bf16), _p(d_sf), None) == 0
    got_dnet = (d_net.astype(np.uint32) << 16).view(np.float32) if bf16 else d_net
    assert np.abs(got_dnet - net.grad.numpy()).max() <= (8e-3 if bf16 else 2e-6) * np.abs(net.grad.numpy()).max()
    assert np.abs(d_sf - sf.grad.numpy().reshape(rows, C)).max() <= 2e-6 * np.abs(sf.grad.numpy()).max()
    # missing upstream gradients are zeros
    assert emu.p2r_vote_tail_grad(None, None, _p(feat), _p(norm), rows, C, _p(d_net), int(bf16), _p(d_sf), None) == 0
    assert not d_sf.any() and not d_net.any()


def test_no_shared_memory_race_under_thread_sanitizer(tmp_path):
    """The emulator's threads are real threads, so a missing __syncthreads (or an unordered global hand-over in the
    last-block pattern) is a data race ThreadSanitizer can see.  Instrumented build, one pass over every kernel: no race.
    A mutant with the barrier after the shared-memory staging of detection_loss_kernel removed must be reported -- the
    check that the instrument is actually looking."""
    import os
    import shutil
    import sys
    tsan = subprocess.run(["gcc", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    if not osp.isabs(tsan) or not osp.exists(tsan):
        pytest.skip("libtsan not available")
    inc = [osp.join(ROOT, "include"), osp.join(ROOT, "tests", "csrc")]

    def build_and_run(csrc_dir, tag):
        so = str(tmp_path / ("emu_tsan_%s.so" % tag))
        subprocess.run(["g++", "-O1", "-g", "-fsanitize=thread", "-ffp-contract=off", "-pthread", "-shared", "-fPIC",
                        "-std=c++17", "-w", "-DP2R_HOST_EMULATION", "-DP2R_SM_COUNT=1", "-I", csrc_dir] +
                       sum((["-I", i] for i in inc), []) + [osp.join(csrc_dir, "kernels_emu_entry.cpp"), "-o", so], check=True)
        env = dict(os.environ, LD_PRELOAD=tsan, TSAN_OPTIONS="report_signal_unsafe=0 exitcode=0")
        r = subprocess.run([sys.executable, osp.join(ROOT, "tests", "emu_tsan_driver.py"), so], env=env, capture_output=True,
                           text=True, timeout=600, cwd=ROOT)
        assert "TSAN-DRIVER-DONE" in r.stdout, r.stderr[-2000:]
        races = r.stderr.count("WARNING: ThreadSanitizer: data race")
        if tag == "clean":      # + the streaming BatchNorm kernels (mbarrier rings wrapping, column-sum variant included)
            r = subprocess.run([sys.executable, osp.join(ROOT, "tests", "test_stream_bn_emulated.py"), so], env=env,
                               capture_output=True, text=True, timeout=600, cwd=ROOT)
            assert "SBN-DRIVER-DONE" in r.stdout, r.stderr[-2000:]
            races += r.stderr.count("WARNING: ThreadSanitizer: data race")
        return races

    def stage(tag, mutate):
        d = tmp_path / tag
        d.mkdir()
        src = osp.join(ROOT, "pose2room_b200", "csrc")
        entry = open(osp.join(ROOT, "tests", "csrc", "kernels_emu.cpp")).read()
        emulated = [l.split("/")[-1].rstrip('"\n') for l in entry.splitlines() if l.startswith('#include "../../pose2room_b200/csrc/')]
        for f in os.listdir(src):
            if f.endswith((".h", ".cuh")) or f in emulated:
                shutil.copy(osp.join(src, f), d / f)
        entry = entry.replace("../../pose2room_b200/csrc/", "")
        (d / "kernels_emu_entry.cpp").write_text(entry)
        if mutate:
            p = d / "loss_ops.cu"
            s = p.read_text()
            barrier = "s_center[i] = __ldg(a.center + (size_t)b * P * 3 + i);\n    __syncthreads();\n"
            assert barrier in s
            p.write_text(s.replace(barrier, barrier.replace("    __syncthreads();\n", "")))
        return str(d)

    assert build_and_run(stage("clean", False), "clean") == 0
    assert build_and_run(stage("mutant", True), "mutant") > 0


@pytest.mark.parametrize("variant", [1, 2])
def test_make_batch_kernels_under_emulation_equal_the_host_harness(emu_dl, tmp_path_factory, variant):
    """Both data-movement variants of the sample -> batch kernel (csrc/dataloader_ops.cu, launchers included) against the
    sequential harness over the same arithmetic header (tests/csrc/make_batch_host.c, itself bit-exact against the
    reference goldens in test_dataloader_math.py): ragged samples, up- and down-sampling, augmented and plain items, a tail
    group (T = 513), 4-channel output -- and, for variant 2, several work items per persistent CTA (its 3-stage ring wraps)."""
    from tests import test_dataloader_math as TD
    host = TD.host_lib.__wrapped__(tmp_path_factory)
    rng = np.random.default_rng(5)
    J, B, T, C = 25, 5, 513, 4
    frames = np.array([30, 700, 1, 513, 1200])
    fs = np.zeros(len(frames) + 1, np.int64)
    fs[1:] = np.cumsum(frames)
    joints = rng.standard_normal((int(fs[-1]), J, 3)).astype(np.float32)
    votes = rng.standard_normal((int(fs[-1]), J, 10)).astype(np.float32)
    votes[..., 0] = rng.integers(0, 2, size=votes.shape[:2])
    ids = np.array([4, 0, 2, 3, 1], np.int32)
    params = np.zeros((B, 16))
    for b in range(B):
        if b % 2 == 0:                                      # augmented: flip / rotation / shift / floor
            th = rng.uniform(-np.pi, np.pi)
            params[b, 0], params[b, 1] = 1.0, float(b % 4 == 0)
            params[b, 2:11] = np.array([[np.cos(th), 0, -np.sin(th)], [0, 1, 0], [np.sin(th), 0, np.cos(th)]]).ravel()
            params[b, 11:14] = [0.3, 0.0, -0.2]
        params[b, 14] = 0.05
    want = [np.empty((B, T, J, C), np.float32), np.empty((B, T, J, 9), np.float32), np.empty((B, T, J), np.int64)]
    host.host_make_batch(_vp(joints), _vp(votes), _vp(fs), _vp(ids), _vp(params), B, T, J, C, *[_vp(a) for a in want])
    got = [np.full((B, T, J, C), np.nan, np.float32), np.full((B, T, J, 9), np.nan, np.float32), np.full((B, T, J), -5, np.int64)]
    rc = emu_dl.p2r_make_batch_variant(variant, _p(joints), _p(votes), _p(fs), _p(ids), _p(params), B, T, J, C,
                                       *[_p(a) for a in got], None)
    assert rc == 0, emu_dl.emu_last_error()
    for w, g_ in zip(want, got):
        assert np.array_equal(w.view(np.uint8), g_.view(np.uint8))      # bit-exact, every element written


@pytest.fixture(scope="module")
def emu_dl(emu):
    fn = emu.p2r_make_batch_variant
    fn.argtypes = _lib.SIGNATURES["p2r_make_batch_variant"]
    fn.restype = ctypes.c_int
    return emu
