"""Which lines of this package launch the torch glue kernels (at::native::*) of one train step?

One EAGER step of bench.py's workload under torch.profiler with Python stacks; every aten op that spent device time is
attributed to the innermost frame inside pose2room_b200/ (or bench.py) and the table is sorted by device time.  The step
that is timed is the captured graph; this tool only says where its non-library launches come from.

    python tools/glue_trace.py [--batch 32] > gpurun_out/glue_trace.txt
"""
import argparse
import collections
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--top", type=int, default=60)
    ap.add_argument("--gemms", action="store_true", help="instead: every GEMM of the step timed alone (ops.PROFILE), by shape")
    args = ap.parse_args()
    from pose2room_b200 import gemm_sm100, ops, synthetic
    from pose2room_b200.config import P2RConfig
    from pose2room_b200.p2rnet import P2RNet
    from torch.profiler import ProfilerActivity, profile
    dev = torch.device("cuda:0")
    gemm_sm100.install()
    torch.manual_seed(42)
    np.random.seed(42)
    net = P2RNet(P2RConfig(mode="train", joint_num=25, num_frames=1024, precision="bf16"))
    net.load_state_dict(synthetic.deterministic_state_dict(net.state_dict(), seed=7))
    net = net.to(dev).train()
    params = [p for p in net.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-3, fused=True, capturable=True)
    ops.register_weight_shadows(net)
    host = synthetic.make_batch(args.batch, 1024, 25, seed=1234)
    data = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in host.items()}

    def step():
        opt.zero_grad(set_to_none=True)
        with ops.overlap_weight_grads():
            ep = net(data)
            loss = net.loss(ep, data)["total"]
            loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()

    if args.gemms:
        ops.PROFILE["log"], ops.PROFILE["on"] = [], True
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ops.PROFILE["on"] = False
        tab = collections.defaultdict(list)
        for r in ops.PROFILE["log"]:
            tab[r[:4]].append(r[4].elapsed_time(r[5]) * 1e3)
        print("%-10s %9s %6s %6s %5s %9s %9s %9s" % ("tag", "M", "N", "K", "n/step", "us", "TFLOP/s", "GB/s(min)"))
        for (tag, m, n, k), us in sorted(tab.items(), key=lambda kv: -sum(kv[1])):
            t = sum(us) / len(us)
            flops = 2.0 * m * n * k
            byts = 2.0 * (m * k + m * n) + 4.0 * n * k      # the two tall operands in bf16 + the small one
            print("%-10s %9d %6d %6d %5d %9.1f %9.1f %9.1f" % (tag, m, n, k, len(us) // 3, t, flops / t / 1e6, byts / t / 1e3))
        return

    # (a) who calls them: a dispatch mode sees every aten op of the step (forward, and the Python backward functions on
    # the autograd thread) with the Python stack that issued it; bytes written stand in for time.
    from torch.utils._python_dispatch import TorchDispatchMode
    import traceback
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    calls = collections.defaultdict(lambda: [0, 0])

    class Trace(TorchDispatchMode):
        def __torch_dispatch__(self, func, types, args=(), kwargs=None):
            out = func(*args, **(kwargs or {}))
            name = str(func)
            if any(name.startswith("aten." + k) for k in ("view", "reshape", "_unsafe_view", "expand", "permute", "transpose", "t.",
                                                          "slice", "select", "unsqueeze", "squeeze", "detach", "alias", "as_strided",
                                                          "empty", "_local_scalar", "unbind", "split", "is_", "sym_", "stride", "size")):
                return out
            where = "(no package frame: autograd built-in)"
            for fr in reversed(traceback.extract_stack()[:-1]):
                if "pose2room_b200/" in fr.filename or fr.filename.endswith("bench.py"):
                    where = "%s:%d" % (fr.filename.replace(root + "/", ""), fr.lineno)
                    break
            nbytes, shape = 0, ""
            for o in (out if isinstance(out, (tuple, list)) else (out,)):
                if isinstance(o, torch.Tensor):
                    nbytes += o.numel() * o.element_size()
                    shape = shape or "%s %s" % (tuple(o.shape), str(o.dtype).replace("torch.", ""))
            if where.startswith("(no package"):
                where += " " + shape
            c = calls[(name, where)]
            c[0] += 1
            c[1] += nbytes
            return out

    with Trace():
        step()
    torch.cuda.synchronize()
    rows = sorted(calls.items(), key=lambda kv: -kv[1][1])
    print("aten calls of one eager step by call site: %d calls at %d sites" % (sum(v[0] for _, v in rows), len(rows)))
    for (name, where), (n, nb) in rows[:args.top]:
        print("%10.2f MB %4d x  %-34s %s" % (nb / 1e6, n, name, where))
    by_count = sorted(calls.items(), key=lambda kv: -kv[1][0])
    print("\nby call count:")
    for (name, where), (n, nb) in by_count[:args.top]:
        print("%10.2f MB %4d x  %-34s %s" % (nb / 1e6, n, name, where))

    # (b) what they cost on the device
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    table = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        dt = getattr(ev, "self_device_time_total", 0) or 0
        if dt <= 0 or not ev.name.startswith("aten::"):
            continue
        table[ev.name][0] += 1
        table[ev.name][1] += dt
    rows = sorted(table.items(), key=lambda kv: -kv[1][1])
    print("\naten ops with device time in one eager step: %d calls, %.1f us" % (sum(v[0] for _, v in rows), sum(v[1] for _, v in rows)))
    for name, (n, us) in rows[:30]:
        print("%8.1f us %4d x  %s" % (us, n, name))


if __name__ == "__main__":
    main()
