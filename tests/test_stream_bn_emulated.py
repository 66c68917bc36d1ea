"""CPU: the bulk-TMA streaming BatchNorm KERNELS (csrc/stream_bn.cu, unmodified: producer lane + 8 consumer warps, a ring
of full / empty mbarriers per operand, persistent CTAs) executed by the host emulator with emulated mbarriers and bulk
copies (tests/csrc/cuda_emu.h), against numpy: forward statistics, affine (+ residual, ReLU, 1-bit mask), backward
statistics in every ReLU-mask mode, backward apply incl. the per-(row % period) column sums of dx.  Built with
P2R_SM_COUNT=1 so that two persistent CTAs share ~40 tiles and every ring wraps several times.

Why it exists: the column-sum variant (P2R_FUSED_COLSUM=1) is off because bench warm-up steps stalled with it twice
(DESIGN.md section 3).  Here the kernel's own protocol is run to completion -- a wait that makes no progress aborts with a
diagnostic -- and test_kernels_emulated.py's sanitizer pass covers it for races: it completes, is correct and race-free, so
the stall is not in the ring protocol (look at the stream interplay next)."""
import ctypes
import os.path as osp
import subprocess

import numpy as np
import pytest

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
C = 64


def build(path, extra=()):
    inc = [osp.join(ROOT, "include"), osp.join(ROOT, "pose2room_b200", "csrc"), osp.join(ROOT, "tests", "csrc")]
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-pthread", "-shared", "-fPIC", "-std=c++17", "-w",
                    "-DP2R_HOST_EMULATION", "-DP2R_SM_COUNT=1"] + list(extra) + sum((["-I", i] for i in inc), []) +
                   [osp.join(ROOT, "tests", "csrc", "kernels_emu.cpp"), "-o", path], check=True)


def bind(path):
    lib = ctypes.CDLL(path)
    vp, ll, ci = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
    lib.emu_stream_col_stats.argtypes = [vp, ll, vp, vp]
    lib.emu_stream_col_bwd_stats.argtypes = [vp, vp, vp, ll, vp, vp, ci, vp, vp, vp, vp]
    lib.emu_stream_affine_act.argtypes = [vp, ll, vp, vp, vp, ci, vp, vp]
    lib.emu_stream_bn_bwd_apply.argtypes = [vp, vp, vp, ll, vp, vp, vp, vp, vp, ci, vp, vp, vp, vp, ci]
    lib.emu_stream_colsum_period.argtypes = [vp, ll, ci, vp]
    return lib


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu_sbn") / "kernels_emu_sm1.so")
    build(so)
    return bind(so)


def to_bf16(a):
    u = np.ascontiguousarray(a, np.float32).view(np.uint32)
    return ((u + 0x7fff + ((u >> 16) & 1)) >> 16).astype(np.uint16)


def from_bf16(u):
    return (u.astype(np.uint32) << 16).view(np.float32)


def _p(a):
    return a.ctypes.data if a is not None else None


def run_all(lib, seed=0):
    """One pass over every mode; returns the worst relative errors (asserted by the caller)."""
    rng = np.random.default_rng(seed)
    worst = {}

    def note(k, v):
        worst[k] = max(worst.get(k, 0.0), float(v))
    for M, relu, period in [(64 * 40 + 17, 2, 25), (64 * 37, 1, 25), (64 * 33 + 5, 3, 7), (64 * 20, 0, 0)]:
        x16, dy16 = to_bf16(rng.normal(size=(M, C))), to_bf16(rng.normal(size=(M, C)))
        x, dy = from_bf16(x16), from_bf16(dy16)
        mean, rstd = (0.1 * rng.normal(size=C)).astype(np.float32), (1 + 0.2 * rng.random(C)).astype(np.float32)
        gamma, beta = (1 + 0.1 * rng.normal(size=C)).astype(np.float32), (0.1 * rng.normal(size=C)).astype(np.float32)
        scale, shift = (gamma * rstd).astype(np.float32), (beta - mean * gamma * rstd).astype(np.float32)
        # forward statistics
        s1, s2 = np.zeros(C), np.zeros(C)
        assert lib.emu_stream_col_stats(_p(x16), M, _p(s1), _p(s2)) == 0
        note("fwd_stats", max(np.abs(s1 - x.astype(np.float64).sum(0)).max() / np.abs(x).sum(0).max(),
                              np.abs(s2 - (x.astype(np.float64) ** 2).sum(0)).max() / (x.astype(np.float64) ** 2).sum(0).max()))
        # affine (+ residual + ReLU + mask)
        res16 = to_bf16(rng.normal(size=(M, C)))
        y16, mask_out = np.zeros((M, C), np.uint16), np.zeros((M, 8), np.uint8)
        assert lib.emu_stream_affine_act(_p(x16), M, _p(scale), _p(shift), _p(res16), 1, _p(y16), _p(mask_out)) == 0
        want_y = np.maximum(x * scale + shift + from_bf16(res16), 0)
        note("affine", np.abs(from_bf16(y16) - want_y).max() / np.abs(want_y).max())
        assert np.array_equal(np.unpackbits(mask_out, axis=1, bitorder="little").reshape(M, C) > 0, from_bf16(y16) > 0)
        # backward: the ReLU mask in each of its four encodings
        if relu == 1:
            mask, yarg = from_bf16(y16) > 0, y16
        elif relu == 2:
            mask, yarg = (x * scale + shift) > 0, None
        elif relu == 3:
            mask = rng.random((M, C)) > 0.4
            yarg = np.packbits(mask.reshape(M, 8, 8), axis=2, bitorder="little").reshape(M, 8).copy()
        else:
            mask, yarg = np.ones((M, C), bool), None
        dz = np.where(mask, dy, 0).astype(np.float64)
        xhat = (x.astype(np.float64) - mean) * rstd
        b1, b2 = np.zeros(C), np.zeros(C)
        assert lib.emu_stream_col_bwd_stats(_p(dy16), _p(x16), _p(yarg), M, _p(mean), _p(rstd), relu, _p(b1), _p(b2),
                                            _p(scale), _p(shift)) == 0
        note("bwd_stats", max(np.abs(b1 - dz.sum(0)).max() / np.abs(dz.sum(0)).max(),
                              np.abs(b2 - (dz * xhat).sum(0)).max() / np.abs((dz * xhat).sum(0)).max()))
        dx16, dres16 = np.zeros((M, C), np.uint16), np.zeros((M, C), np.uint16)
        cs = np.zeros((max(period, 1), C))
        assert lib.emu_stream_bn_bwd_apply(_p(dy16), _p(x16), _p(yarg), M, _p(mean), _p(rstd), _p(scale), _p(b1), _p(b2), relu,
                                           _p(dx16), _p(dres16), _p(shift), _p(cs) if period else None, period) == 0
        want = scale * (dz - b1 / M - xhat * b2 / M)
        got = from_bf16(dx16)
        note("dx", np.abs(got - want).max() / np.abs(want).max())
        note("dres", np.abs(from_bf16(dres16) - dz).max())
        if period:
            wcs = np.zeros((period, C))
            np.add.at(wcs, np.arange(M) % period, got.astype(np.float64))
            note("colsum", np.abs(cs - wcs).max() / np.abs(wcs).max())
    return worst


def test_streaming_batchnorm_kernels_under_emulation(emu):
    w = run_all(emu)
    assert w["fwd_stats"] < 1e-6 and w["bwd_stats"] < 1e-6 and w["colsum"] < 1e-6 and w["dres"] == 0.0
    assert w["affine"] < 8e-3 and w["dx"] < 8e-3                     # outputs are rounded to bf16


@pytest.mark.parametrize("period,frames", [(25, 41), (25, 200), (32, 33), (2, 257), (13, 64)])
def test_periodic_column_sums_under_emulation(emu, period, frames):
    """stream_colsum_period_kernel (the graph convolution's bias gradient: rows of 64 channels cycling through `period`
    joints, tiles of whole periods, register-resident sums): equal to a float64 sum per (joint, channel), ring wrapping,
    ragged last tile, canary behind the output intact."""
    rng = np.random.default_rng(period * 1000 + frames)
    rows = period * frames
    x = to_bf16(rng.normal(0.0, 1.0, size=(rows, 64)).astype(np.float32))
    out = np.zeros(period * 64 + 8, np.float64)
    out[-8:] = 123.0
    assert emu.emu_stream_colsum_period(_p(x), rows, period, _p(out)) == 0
    want = from_bf16(x).astype(np.float64).reshape(frames, period, 64).sum(0).reshape(-1)
    assert np.all(out[-8:] == 123.0)
    assert np.abs(out[:-8] - want).max() <= 2e-4 * max(1.0, np.abs(want).max())       # fp32 partial sums per thread


if __name__ == "__main__":      # sanitizer driver: python tests/test_stream_bn_emulated.py <instrumented .so>
    import sys
    print(run_all(bind(sys.argv[1])))
    print("SBN-DRIVER-DONE")
