"""Operator microbench (not the bench.py contract): times each native op with CUDA events on the live and the
large-cloud shapes, next to the unmodified reference kernels from oracle/_ref when present.
Usage (GPU box): python microbench.py > gpurun_out/microbench.json"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (run as `python tools/<name>.py`)

import json
import sys

import torch

from pose2room_b200 import ext, synthetic


def timeit(fn, iters=20, warm=3):
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
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ref = None
    try:
        from oracle import build_ref_ext
        ref = build_ref_ext.load_ref_ext()
    except Exception as e:  # noqa
        print("no reference ext:", e, file=sys.stderr)
    dev = torch.device("cuda:0")
    out = []
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
            rec = dict(shape=name, impl=iname,
                       fps_ms=timeit(lambda: mod.furthest_point_sampling(xyz, M), iters=5 if N > 1000 else 20),
                       ball_query_ms=timeit(lambda: mod.ball_query(new_xyz, xyz, r, ns)),
                       group_ms=timeit(lambda: mod.group_points(feats, bq)),
                       group_grad_ms=timeit(lambda: mod.group_points_grad(go, bq, N)),
                       three_nn_ms=timeit(lambda: mod.three_nn(xyz, new_xyz)))
            out.append(rec)
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
