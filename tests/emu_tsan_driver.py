"""Run by tests/test_kernels_emulated.py in a process of its own with libtsan preloaded: one pass of every emulated kernel
(ThreadSanitizer-instrumented build of tests/csrc/kernels_emu.cpp given as argv[1]) on shapes with tail blocks."""
import ctypes
import os.path as osp
import sys

import numpy as np
import torch

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
from pose2room_b200 import _lib  # noqa: E402
from tests import test_kernels_emulated as TE  # noqa: E402
from tests import test_loss_math as TL  # noqa: E402

lib = ctypes.CDLL(sys.argv[1])
for name in ("p2r_detection_loss", "p2r_detection_loss_grad", "p2r_gmm_mix", "p2r_gmm_mix_grad",
             "p2r_detection_loss_workspace", "p2r_gmm_mix_workspace", "p2r_vote_tail", "p2r_vote_tail_grad"):
    fn = getattr(lib, name)
    fn.argtypes = _lib.SIGNATURES[name]
    fn.restype = _lib._RESTYPES.get(name, ctypes.c_int)
lib.emu_last_error.restype = ctypes.c_char_p
est, gt, so = TL.make_case(11, B=2, T=300, S=200, P=150, heading_dtype=torch.float32)
TE.emulated_loss(lib, est, gt, so)
rows, G, D = 77, 33, 3
rng = np.random.default_rng(0)
lg, mu = rng.normal(size=(rows, G)).astype(np.float32), rng.normal(size=(G, D)).astype(np.float32)
ls, eps = np.zeros((G, D), np.float32), rng.normal(size=(rows, G, 1, D)).astype(np.float32)
dout, out = rng.normal(size=(rows, D)).astype(np.float32), np.zeros((rows, D), np.float32)
p = lambda a: a.ctypes.data
assert lib.p2r_gmm_mix(p(lg), 0, p(mu), 0, p(ls), p(eps), rows, G, D, p(out), None) == 0
n = lib.p2r_gmm_mix_workspace(rows, G, D)
ws, dl = np.zeros(n), np.zeros((rows, G), np.float32)
dmu, dls = np.zeros((G, D), np.float32), np.zeros((G, D), np.float32)
assert lib.p2r_gmm_mix_grad(p(lg), 0, p(mu), 0, p(ls), p(eps), p(dout), rows, G, D, p(dl), p(dmu), p(dls), p(ws), n, None) == 0
C = 256
net, sf = rng.normal(size=(rows, 3 + C)).astype(np.float32), rng.normal(size=(rows, C)).astype(np.float32)
xyz_in, xyz, feat, norm = rng.normal(size=(rows, 3)).astype(np.float32), np.zeros((rows, 3), np.float32), \
    np.zeros((rows, C), np.float32), np.zeros(rows, np.float32)
assert lib.p2r_vote_tail(p(net), 0, p(xyz_in), 3, p(sf), rows, C, p(xyz), p(feat), p(norm), None) == 0
dn, ds = np.zeros_like(net), np.zeros_like(sf)
assert lib.p2r_vote_tail_grad(p(xyz), p(feat), p(feat), p(norm), rows, C, p(dn), 0, p(ds), None) == 0
# sample -> batch, both variants (variant 2: several work items per CTA, so its cp.async ring wraps)
fn = lib.p2r_make_batch_variant
fn.argtypes = _lib.SIGNATURES["p2r_make_batch_variant"]
fn.restype = ctypes.c_int
Jn, Bn, Tn = 7, 2, 81       # 22 work items; the sanitizer build has P2R_SM_COUNT=1 -> one persistent CTA walks all of them
fs = np.array([0, 50, 61], np.int64)
jt, vt = rng.normal(size=(61, Jn, 3)).astype(np.float32), rng.normal(size=(61, Jn, 10)).astype(np.float32)
ids, par = np.array([1, 0], np.int32), np.zeros((Bn, 16))
o = [np.zeros((Bn, Tn, Jn, 3), np.float32), np.zeros((Bn, Tn, Jn, 9), np.float32), np.zeros((Bn, Tn, Jn), np.int64)]
for variant in (1, 2):
    assert fn(variant, p(jt), p(vt), p(fs), p(ids), p(par), Bn, Tn, Jn, 3, p(o[0]), p(o[1]), p(o[2]), None) == 0
print("TSAN-DRIVER-DONE")
