"""Kernel timeline of ONE replay of the captured train-step graph (torch.profiler / CUPTI): busy vs idle time,
per-stream time, largest gaps.  Diagnostic only -- not a bench number."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (run as `python tools/<name>.py`)

import json
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

from pose2room_b200 import gemm_sm100, ops, synthetic
from pose2room_b200.config import P2RConfig
from pose2room_b200.p2rnet import P2RNet

dev = torch.device("cuda:0")
gemm_sm100.install()
torch.manual_seed(42)
np.random.seed(42)
net = P2RNet(P2RConfig(mode="train", joint_num=25, num_frames=1024, precision="bf16"))
net.load_state_dict(synthetic.deterministic_state_dict(net.state_dict(), seed=7))
net = net.to(dev).train()
params = [p for p in net.parameters() if p.requires_grad]
opt = torch.optim.AdamW(params, lr=1e-3, fused=True, capturable=True)
data = {k: v.to(dev) for k, v in synthetic.make_batch(32, 1024, 25, seed=1234).items() if isinstance(v, torch.Tensor)}


def step():
    opt.zero_grad(set_to_none=True)
    with ops.overlap_weight_grads():
        ep = net(data)
        loss = net.loss(ep, data)["total"]
        loss.backward()
    opt.step()
    return loss


side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
opt.zero_grad(set_to_none=True)
with torch.cuda.graph(g):
    step()
torch.cuda.synchronize()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.replay()
    torch.cuda.synchronize()
prof.export_chrome_trace("gpurun_out/graph_trace.json")
allev = json.load(open("gpurun_out/graph_trace.json"))["traceEvents"]
gpu_other = [e for e in allev if e.get("cat") in ("gpu_memcpy", "gpu_memset") and "ts" in e]
print("gpu memcpy/memset events:", len(gpu_other), "total us %.1f" % sum(e["dur"] for e in gpu_other))
for e in sorted(gpu_other, key=lambda e: -e["dur"])[:8]:
    print("   %s dur %.1f us args %s" % (e["name"][:40], e["dur"], {k: e["args"][k] for k in list(e["args"])[:6]}))
cats = {}
for e in allev:
    cats[e.get("cat")] = cats.get(e.get("cat"), 0) + 1
print("event categories:", cats)
ev = [e for e in allev if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
t1 = max(e["ts"] + e["dur"] for e in ev)
print("kernels %d, span %.1f us, sum of kernel time %.1f us" % (len(ev), t1 - t0, sum(e["dur"] for e in ev)))
# busy time (union of intervals)
busy, cur_s, cur_e = 0.0, None, None
gaps = []
for e in ev:
    s, d = e["ts"], e["ts"] + e["dur"]
    if cur_e is None:
        cur_s, cur_e = s, d
    elif s <= cur_e:
        cur_e = max(cur_e, d)
    else:
        busy += cur_e - cur_s
        gaps.append((s - cur_e, cur_e - t0, e["name"][:60]))
        cur_s, cur_e = s, d
busy += cur_e - cur_s
print("GPU busy (union) %.1f us, idle %.1f us in %d gaps" % (busy, (t1 - t0) - busy, len(gaps)))
streams = {}
for e in ev:
    st = e["args"].get("stream")
    streams[st] = streams.get(st, 0.0) + e["dur"]
print("per-stream kernel time:", {k: round(v, 1) for k, v in streams.items()})
# idle time per 500-us window
win = 500.0
nw = int((t1 - t0) / win) + 1
idle_w = [0.0] * nw
for gdur, at, _ in gaps:
    idle_w[int(at / win)] += gdur
print("idle us per 500-us window:", [round(x) for x in idle_w])
cnt_w = [0] * nw
for e in ev:
    cnt_w[int((e["ts"] - t0) / win)] += 1
print("kernels per 500-us window:", cnt_w)
print("largest gaps:")
for gdur, at, nm in sorted(gaps, reverse=True)[:12]:
    print("  %.1f us at t=%.0f before %s" % (gdur, at, nm))
# ---- the vote / head / loss region: the windows with > 60 kernels per 500 us
dense = [i for i, c in enumerate(cnt_w) if c > 60 and i > 0]
if dense:
    lo, hi = dense[0] * win, (dense[-1] + 1) * win
    reg = [e for e in ev if lo <= e["ts"] - t0 < hi]
    print("small-kernel region %.0f..%.0f us: %d kernels, sum %.0f us" % (lo, hi, len(reg), sum(e["dur"] for e in reg)))
    agg = {}
    for e in reg:
        k = e["name"][:70]
        a = agg.get(k, [0.0, 0])
        a[0] += e["dur"]; a[1] += 1
        agg[k] = a
    for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        print("  %7.1f us x%-3d %s" % (t, c, k))
    st = {}
    for e in reg:
        st[e["args"].get("stream")] = st.get(e["args"].get("stream"), 0.0) + e["dur"]
    print("  per-stream in region:", {k: round(v) for k, v in st.items()})
