#!/usr/bin/env python
"""bench.py -- headline benchmark of the P2RNet hot path on B200 (contract in the task brief / DESIGN.md).

  python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K --warmup W   # reference arm: the reference algorithm on host CPU

Workload (BASELINE.json metric, config #3 per GPU): synthetic pose sequences, B = 32 per GPU, T = 1024 frames,
J = 25 joints; one "step" = P2RNet forward + detection loss + backward + AdamW update (+ gradient all-reduce for
N > 1) on one batch.  `value` = sequences/s with the batch resident in HBM; `e2e` = the same through the public
model API with the batch in pinned host memory (H2D inside the timed region, loss read back each step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

B_PER_GPU, T_FRAMES, JOINTS = 32, 1024, 25
# DRAM bytes of ONE launch of the dominant kernel (graph-conv forward GEMM, block-sparse, fused statistics) from the
# ncu --set full capture of round 2: dram__bytes_read.sum 109.07 MB + dram__bytes_write.sum 66.73 MB
# (profiles/r02_ncu_hot_kernels_summary.txt; algorithmic bytes: X 104.9 MB + W_eff 5.1 MB + Y 104.9 MB -- part of Y is still
# in the 126 MB L2 when the kernel ends); the weight-gradient GEMM: 214.4 MB (its operands once: 209.7 MB + dW 10 MB)
GCN_FWD_DRAM_BYTES_NCU = 109.065e6 + 66.726e6
GCN_DW_DRAM_BYTES_NCU = 214.43e6
GCN_DW_PAIR_DRAM_BYTES_NCU = 231.39e6 + 14.04e6
# algorithmic (conv 64->704 + einsum, the reference's formulation) forward FLOPs of ONE graph convolution for ONE
# sequence at T=1024, J=25: (13.84 + 5.41) GFLOP / 6 blocks  (SURVEY.md section 8d / BASELINE.md section 3)
GCN_ALGO_GFLOP_PER_SEQ = (13.84 + 5.41) / 6.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("P2R_PRECISION", "auto"), choices=["auto", "bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--data-path-variants", action="store_true", help="internal: A/B of the make_batch kernel variants")
    ap.add_argument("--leg", default=None, choices=["gpu_reference", "extras"],
                    help="internal: one of the secondary legs of bench_legs.py (run by the supervisor in a child process)")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed regions (B200_PROFILING.md).  NVML is polled every 10 ms from a
    thread (a 10-step timed region lasts ~0.1 s: spawning nvidia-smi per sample would catch one sample at best);
    `timed(True/False)` brackets the regions whose samples are summarised.  Falls back to nvidia-smi without pynvml."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.in_region = index, [], False, False
        self.nvml, self.handle, self.max_mhz = None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            try:     # CUDA ordinal -> the same physical GPU in NVML (CUDA_VISIBLE_DEVICES may renumber)
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid).replace("GPU-", "")
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def timed(self, on):
        self.in_region = on

    def _sample_nvml(self):
        n = self.nvml
        mhz = int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        return mhz, mask

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        f = [x.strip() for x in out.split(",")]
        if self.max_mhz is None and f[1].isdigit():
            self.max_mhz = int(f[1])
        mask = 0
        for bit, col in ((0x8, 2), (0x40, 3), (0x20, 4), (0x4, 5)):
            if f[col].lower().startswith("active"):
                mask |= bit
        return int(f[0]), mask

    def run(self):
        while not self.stop_flag:
            try:
                mhz, mask = self._sample_nvml() if self.nvml is not None else self._sample_smi()
                self.rows.append((self.in_region, mhz, mask))
            except Exception:
                pass
            time.sleep(0.01 if self.nvml is not None else 0.2)

    def summary(self):
        rows = [r for r in self.rows if r[0]] or self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        sm = sorted(r[1] for r in rows)
        mask = 0
        for r in rows:
            mask |= r[2]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": [name for bit, name in self.REASONS.items() if mask & bit],
                "samples": len(rows), "samples_in_timed_regions": sum(1 for r in self.rows if r[0]),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------ reference arm (CPU)
def build_cpu_reference(batch):
    """The reference algorithm on host cores: oracle/model_ref.py (plain fp32 PyTorch port of the reference's
    modules, validated against the real reference in the build container) + oracle/pointnet2_ref.c."""
    from oracle.model_ref import RefP2RNet
    from pose2room_b200 import synthetic
    from pose2room_b200.config import P2RConfig
    from pose2room_b200.p2rnet import P2RNet
    torch.manual_seed(42)
    np.random.seed(42)
    template = P2RNet(P2RConfig(mode="train", joint_num=JOINTS, num_frames=T_FRAMES)).state_dict()
    sd = synthetic.deterministic_state_dict(template, seed=7)
    net = RefP2RNet(sd, joint_num=JOINTS, num_seeds=512, num_target=128, training=True)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3)
    data = synthetic.make_batch(batch, T_FRAMES, JOINTS, seed=1234)

    def step():
        opt.zero_grad(set_to_none=True)
        ep = net.forward(data)
        loss = net.loss(ep, data)["total"]
        loss.backward()
        opt.step()
        return float(loss)
    return step


def time_cpu_reference(batch, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    step = build_cpu_reference(batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 2  # bounded sample of the workload (memory + minutes): per-sequence cost is batch-independent on CPU
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    seqs, per_step = time_cpu_reference(batch, steps, warmup)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "pose-sequences/sec fwd+bwd (B=32, T=1024, J=25)", "value": seqs,
        "unit": "sequences/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "P2RNet train step fwd+loss+bwd+AdamW, T=1024, J=25, 512 seeds, 128 proposals",
                   "per_gpu_batch": B_PER_GPU, "sample_batch": batch},
        "cpu_baseline": {"value": seqs, "unit": "sequences/s", "cores": cores, "kind": "port",
                         "sample": "B=%d sequences x %d steps of the same workload (T=1024, J=25) on %d host threads; "
                                   "oracle/model_ref.py + oracle/pointnet2_ref.c" % (batch, steps, cores)},
        "e2e": {"value": seqs, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ supervision
# A bench that hangs helps nobody, so: (1) every measuring process carries a watchdog that exits with an error when no
# phase boundary has been crossed for P2R_BENCH_STALL_S seconds; (2) the single-GPU run is a child process of a small
# supervisor that, if the child stalls, kills it and measures ONCE more -- the SAME configuration (round 1 fell back to a
# single-stream step here; round 2 removed that: the experimental multi-stream configurations that stalled in round 1 ran
# 100 eager steps + 100 replays each without a stall, tools/stress_multistream.py, and a different configuration must
# never stand in for the benched one) -- and says so in `config.retry`.  Multi-rank runs (torchrun) only have the watchdog.
_HEARTBEAT = {"t": time.time(), "phase": "start"}


def beat(phase):
    _HEARTBEAT["t"], _HEARTBEAT["phase"] = time.time(), phase


def start_watchdog():
    limit = float(os.environ.get("P2R_BENCH_STALL_S", "150"))

    def run():
        while True:
            time.sleep(2.0)
            if time.time() - _HEARTBEAT["t"] > _HEARTBEAT.get("limit", limit):
                if _HEARTBEAT.get("exit_code", 17) == 0:     # the line is out, only the teardown is stuck: leave cleanly
                    sys.stderr.write("bench.py: process-group teardown did not return in %.0f s -- exiting\n" % _HEARTBEAT["limit"])
                    sys.stderr.flush()
                    os._exit(0)
                sys.stderr.write("bench.py: no progress for %.0f s in phase '%s' -- aborting\n" % (limit, _HEARTBEAT["phase"]))
                sys.stderr.flush()
                if os.environ.get("P2R_BENCH_GDB"):      # diagnostic: which kernels are resident on the GPU right now
                    try:
                        r = subprocess.run(["cuda-gdb", "-p", str(os.getpid()), "-batch", "-ex", "info cuda kernels"],
                                           capture_output=True, text=True, timeout=90)
                        sys.stderr.write("cuda-gdb:\n" + r.stdout[-6000:] + r.stderr[-1500:] + "\n")
                    except Exception as e:
                        sys.stderr.write("cuda-gdb attach failed: %r\n" % (e,))
                    sys.stderr.flush()
                os._exit(_HEARTBEAT.get("exit_code", 17))
    threading.Thread(target=run, daemon=True).start()


def _variants_subprocess(script):
    """Kernel-variant A/B of the sample -> batch path in a process of its own, bounded in time; never raises."""
    import signal
    if os.environ.get("P2R_BENCH_VARIANTS", "1") == "0":
        return None
    try:
        p = subprocess.Popen([sys.executable, script, "--data-path-variants"], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, start_new_session=True)
        try:
            out, _ = p.communicate(timeout=float(os.environ.get("P2R_BENCH_VARIANTS_TIMEOUT_S", "120")))
        except subprocess.TimeoutExpired:
            os.killpg(p.pid, signal.SIGKILL)
            p.wait()
            return {"error": "timed out"}
        lines = [l for l in out.decode().splitlines() if l.startswith("{")]
        return json.loads(lines[-1]) if lines else {"error": "no output (exit code %s)" % p.returncode}
    except Exception as e:
        return {"error": repr(e)}


# Alternative configurations of the SAME workload, each measured by a fresh child of this script on the same box with the
# same step count after the headline measurement is safe, bounded in time, reported next to the headline -- never instead
# of it.  `fp32_mode` is the precision whose outputs are held to 1e-4 of the reference goldens (tests/test_model_gpu.py);
# the headline is the bf16 mode north_star sanctions for the dense layers.
EXPERIMENTS = [
    ("fp32_mode", {}, ["--precision", "fp32"]),
    ("unfused_loss+gmm+vote", {"P2R_FUSED_LOSS": "0", "P2R_FUSED_GMM": "0", "P2R_FUSED_VOTE": "0"}, []),
]


def _experiments(script, headline):
    import signal
    if os.environ.get("P2R_BENCH_EXPERIMENTS", "1") == "0":
        return None
    deadline = time.time() + float(os.environ.get("P2R_BENCH_EXPERIMENTS_BUDGET_S", "150"))
    per_run = float(os.environ.get("P2R_BENCH_EXPERIMENT_TIMEOUT_S", "75"))
    out = {"baseline": {"ms_per_step": headline.get("ms_per_step"), "first_step_loss": headline.get("first_step_loss"),
                        "kernels_per_step": (headline.get("census") or {}).get("kernels")}}
    for name, extra, extra_args in EXPERIMENTS:
        left = deadline - time.time()
        if left < 40.0:
            out[name] = {"skipped": "time budget of the experiments leg spent"}
            continue
        try:
            env = dict(os.environ, P2R_BENCH_CHILD="1", P2R_BENCH_DATA_PATH="0", **extra)
            env["P2R_BENCH_CENSUS"] = "0" if "--precision" in extra_args else env.get("P2R_BENCH_CENSUS", "1")
            args = [sys.executable, script, "--steps", str(headline.get("steps", 10)), "--warmup",
                    str(headline.get("warmup", 3)), "--no-cpu-baseline"] + extra_args
            p = subprocess.Popen(args, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, start_new_session=True)
            try:
                o, e = p.communicate(timeout=min(per_run, left))
            except subprocess.TimeoutExpired:
                os.killpg(p.pid, signal.SIGKILL)
                p.wait()
                out[name] = {"error": "timed out"}
                continue
            lines = [l for l in o.decode().splitlines() if l.startswith("{")]
            if not lines:       # (a child that fell over in its census leg has printed its complete line before)
                out[name] = {"error": "exit code %s: %s" % (p.returncode, e.decode()[-300:])}
                continue
            d = json.loads(lines[-1])
            out[name] = {"ms_per_step": d.get("ms_per_step"), "value": d.get("value"), "dtype": d.get("dtype"),
                         "first_step_loss": d.get("first_step_loss"), "gpu_launches": d.get("gpu_launches"),
                         "e2e_ms_per_step": (d.get("e2e") or {}).get("ms_per_step"),
                         "cuda_graph": (d.get("config") or {}).get("cuda_graph"),
                         "kernels_per_step": (d.get("census") or {}).get("kernels"),
                         "torch_glue_kernels": (d.get("census") or {}).get("torch_glue_kernels")}
        except Exception as e:
            out[name] = {"error": repr(e)}
    return out


def _leg(script, name, timeout_s):
    """One secondary leg (bench_legs.py) in a process of its own, bounded in time; never raises."""
    import signal
    if os.environ.get("P2R_BENCH_LEGS", "1") == "0":
        return None
    try:
        p = subprocess.Popen([sys.executable, script, "--leg", name], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                             start_new_session=True, env=dict(os.environ, P2R_BENCH_CHILD="1"))
        try:
            o, e = p.communicate(timeout=timeout_s)
        except subprocess.TimeoutExpired:
            os.killpg(p.pid, signal.SIGKILL)
            p.wait()
            return {"error": "timed out after %.0f s" % timeout_s}
        lines = [l for l in o.decode().splitlines() if l.startswith("{")]
        return json.loads(lines[-1]) if lines else {"error": "exit code %s: %s" % (p.returncode, e.decode()[-400:])}
    except Exception as e:
        return {"error": repr(e)}


def supervise(script=None):
    import signal
    script = script or os.path.abspath(__file__)
    attempts = [({}, None), ({}, "second attempt of the same configuration: the first one stalled and was killed")]
    for extra, label in attempts:
        env = dict(os.environ, P2R_BENCH_CHILD="1", **extra)
        p = subprocess.Popen([sys.executable, script] + sys.argv[1:], env=env, stdout=subprocess.PIPE,
                             start_new_session=True)
        try:
            out, _ = p.communicate(timeout=float(os.environ.get("P2R_BENCH_TIMEOUT_S", "420")))
        except subprocess.TimeoutExpired:
            os.killpg(p.pid, signal.SIGKILL)
            out, _ = p.communicate()            # what it had printed before it was killed
            sys.stderr.write("bench.py: attempt timed out\n")
            if not any(l.startswith("{") and '"census": null' in l for l in out.decode().splitlines()):
                continue
        lines = [l for l in out.decode().splitlines() if l.startswith("{")]
        if lines and (p.returncode == 0 or '"census": null' in lines[-1]):
            # (a child that died AFTER printing its complete line -- in the informational census leg -- still counts)
            d = json.loads(lines[-1])
            if p.returncode != 0:
                d["census"] = {"error": "the census leg ended the measuring process (exit code %s)" % p.returncode}
            if label is not None:
                d["config"]["retry"] = label
            if isinstance(d.get("data_path"), dict) and "error" not in d["data_path"]:
                d["data_path"]["variants"] = _variants_subprocess(script)
            if label is None and isinstance(d.get("roofline"), dict) and d.get("n_gpus") == 1:
                d["experiments"] = _experiments(script, d)
                # "the reference on B200" + BASELINE configs #2 (forward only / SA operators) and #5 (eval), bench_legs.py
                d["gpu_reference"] = _leg(script, "gpu_reference", float(os.environ.get("P2R_BENCH_GPU_REF_TIMEOUT_S", "150")))
                ref = d["gpu_reference"] if isinstance(d["gpu_reference"], dict) else {}
                if "fp32" in ref:
                    d["vs_gpu_reference"] = {
                        "what": "this line's value and its fp32_mode sibling over the reference formulation + reference "
                                "kernels on the same B200 (gpu_reference.fp32 / .tf32), same batch, same step",
                        "bf16_over_reference_fp32": d["value"] / ref["fp32"]["train_sequences_per_s"],
                        "bf16_over_reference_tf32": d["value"] / ref["tf32"]["train_sequences_per_s"],
                        "fp32_mode_over_reference_fp32": ((d["experiments"] or {}).get("fp32_mode") or {}).get("value", 0.0) /
                                                         ref["fp32"]["train_sequences_per_s"] or None}
                extras = _leg(script, "extras", float(os.environ.get("P2R_BENCH_EXTRAS_TIMEOUT_S", "150")))
                if isinstance(extras, dict) and "error" in extras and "forward_only" not in extras:
                    d["extras_error"] = extras["error"]
                elif isinstance(extras, dict):
                    d["forward_only"] = extras.get("forward_only")
                    if isinstance(d["forward_only"], dict) and "fp32" in ref:
                        d["forward_only"]["gpu_reference_fp32_forward_ms"] = ref["fp32"]["forward_ms"]
                        d["forward_only"]["gpu_reference_tf32_forward_ms"] = ref["tf32"]["forward_ms"]
                    d["sa_module_forward"] = extras.get("sa_module_forward")
                    d["sa_operators"] = extras.get("sa_operators")
                    d["eval_1k"] = extras.get("eval_1k")
            print(json.dumps(d), flush=True)
            return 0
        if p.returncode != 17:          # a real failure, not a stall: do not hide it behind a retry
            sys.stdout.write(out.decode())
            return p.returncode or 1
        sys.stderr.write("bench.py: attempt stalled (watchdog)\n")
    return 1


# ------------------------------------------------------------------------------------------ sample -> batch leg
def _data_path_store(dev, B):
    """96 synthetic raw samples (random values: the kernel's work does not depend on them) of 1100-1500 frames each,
    packed, plus the train-mode dataset over them."""
    from pose2room_b200 import dataloader as DL
    rng = np.random.default_rng(99)
    n = 3 * B
    frames = rng.integers(1100, 1500, size=n)
    frame_start = np.zeros(n + 1, np.int64)
    frame_start[1:] = np.cumsum(frames)
    F = int(frame_start[-1])
    joints = rng.standard_normal((F, JOINTS, 3), dtype=np.float32)
    votes = rng.standard_normal((F, JOINTS, 10), dtype=np.float32)
    votes[..., 0] = rng.integers(0, 2, size=(F, JOINTS))
    R = np.tile(np.eye(3, dtype=np.float32), (n, 10, 1, 1))
    store = DL.PackedSamples(joints, votes, frame_start, np.full(n, 10, np.int32), np.zeros((n, 10), np.int32),
                             rng.standard_normal((n, 10, 3)).astype(np.float32), R,
                             rng.uniform(0.2, 1.7, (n, 10, 3)).astype(np.float32), ["s%d" % i for i in range(n)])

    class _Cfg:
        config = {"data": {"num_frames": T_FRAMES, "no_height": True, "max_gt_boxes": 10}}
        dataset_config = None
    return store, DL.P2RNet_VirtualHome(_Cfg(), "train", packed=store, device=dev)


def data_path_variants(B=B_PER_GPU, reps=30):
    """A/B of the two data-movement variants of the make_batch kernel (csrc/dataloader_ops.cu) on the data_path_leg
    workload.  Runs in its OWN process (the supervisor starts it after the headline measurement has been printed by the
    measuring child), because variant 2 had not run on a GPU when this was written: whatever it does cannot touch the
    headline numbers.  Prints one JSON object."""
    from pose2room_b200 import _lib, dataloader as DL
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    _lib.load()
    store, ds = _data_path_store(dev, B)
    jd, vd, fsd = store.device_arrays(dev)
    sets = [list(range(k * B, (k + 1) * B)) for k in range(3)]
    draws = [[DL.draw_augmentation() for _ in range(B)] for _ in range(3)]
    params = [torch.from_numpy(ds.host_side(sets[k], draws[k])[0]).to(dev) for k in range(3)]
    ids = [torch.tensor(sets[k], dtype=torch.int32, device=dev) for k in range(3)]
    stream = torch.cuda.current_stream().cuda_stream
    algo_bytes = B * T_FRAMES * JOINTS * (13 * 4 + 12 * 4 + 8)
    peaks = measured_peaks()
    out, ref = {}, None
    for variant in (1, 2):
        try:
            outs = [(torch.empty(B, T_FRAMES, JOINTS, 3, device=dev), torch.empty(B, T_FRAMES, JOINTS, 9, device=dev),
                     torch.empty(B, T_FRAMES, JOINTS, dtype=torch.int64, device=dev)) for _ in range(3)]

            def launch(k):
                _lib.call("p2r_make_batch_variant", variant, jd.data_ptr(), vd.data_ptr(), fsd.data_ptr(), ids[k].data_ptr(),
                          params[k].data_ptr(), B, T_FRAMES, JOINTS, 3, outs[k][0].data_ptr(), outs[k][1].data_ptr(),
                          outs[k][2].data_ptr(), stream)
            for k in range(3):
                launch(k)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(reps):
                launch(i % 3)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            rec = {"kernel_ms": ms, "kernel_GBps": algo_bytes / (ms * 1e-3) / 1e9,
                   "kernel_frac_of_hbm_peak": algo_bytes / (ms * 1e-3) / 1e9 / peaks["hbm"]}
            if ref is None:
                ref = outs
            else:
                rec["bit_identical_to_v1"] = all(torch.equal(a, b) for o, r in zip(outs, ref) for a, b in zip(o, r))
            out["v%d" % variant] = rec
        except Exception as e:
            out["v%d" % variant] = {"error": repr(e)}
            break
    print(json.dumps(out), flush=True)


def data_path_leg(dev, B, reps=30, train_step=None, steps=10):
    """Informational: the sample -> batch path (pose2room_b200/dataloader.py) at the BASELINE shape.
    96 raw samples of 1100-1500 frames resident in HBM; every launch builds a batch of B augmented sequences from a
    different third of them into a different set of output buffers (264 MB rotating footprint > the 126 MB L2).
    `kernel_*` = the p2r_make_batch launch alone (CUDA events), `loader_*` = the public make_batch call including the
    host-side draws, box labels and parameter upload (wall clock around a synchronised loop), `train_from_loader_*` =
    the same train step as the headline fed by that loader (make_batch -> step -> loss read back each step; the
    46 MB-per-step host -> device copy of the contract's `e2e` leg is replaced by 18 KB of parameters and labels)."""
    from pose2room_b200 import _lib, dataloader as DL
    store, ds = _data_path_store(dev, B)
    sets = [list(range(k * B, (k + 1) * B)) for k in range(3)]
    for k in range(3):
        ds.make_batch(sets[k])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(reps):
        ds.make_batch(sets[i % 3])
    torch.cuda.synchronize()
    loader_ms = (time.perf_counter() - t0) / reps * 1e3
    # the launch alone, inputs already on the device
    jd, vd, fsd = store.device_arrays(dev)
    draws = [[DL.draw_augmentation() for _ in range(B)] for _ in range(3)]
    params = [torch.from_numpy(ds.host_side(sets[k], draws[k])[0]).to(dev) for k in range(3)]
    ids = [torch.tensor(sets[k], dtype=torch.int32, device=dev) for k in range(3)]
    outs = [(torch.empty(B, T_FRAMES, JOINTS, 3, device=dev), torch.empty(B, T_FRAMES, JOINTS, 9, device=dev),
             torch.empty(B, T_FRAMES, JOINTS, dtype=torch.int64, device=dev)) for _ in range(3)]
    stream = torch.cuda.current_stream().cuda_stream

    def launch(k):
        _lib.call("p2r_make_batch", jd.data_ptr(), vd.data_ptr(), fsd.data_ptr(), ids[k].data_ptr(), params[k].data_ptr(),
                  B, T_FRAMES, JOINTS, 3, outs[k][0].data_ptr(), outs[k][1].data_ptr(), outs[k][2].data_ptr(), stream)
    for k in range(3):
        launch(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        launch(i % 3)
    e1.record()
    torch.cuda.synchronize()
    kernel_ms = e0.elapsed_time(e1) / reps
    algo_bytes = B * T_FRAMES * JOINTS * (13 * 4 + 12 * 4 + 8)      # 52 B read + 56 B written per (frame, joint)
    peaks = measured_peaks()
    train = {}
    if train_step is not None:
        for i in range(2):
            train_step(ds.make_batch(sets[i % 3])).item()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(steps):
            train_step(ds.make_batch(sets[i % 3])).item()
        g1.record()
        torch.cuda.synchronize()
        t_ms = g0.elapsed_time(g1) / steps
        train = {"train_from_loader_ms_per_step": t_ms, "train_from_loader_sequences_per_s": B / (t_ms * 1e-3)}
    return dict(train, **{"what": "sample -> batch (frame picking + flip/rotate/translate + casts + collate) from HBM-resident raw samples",
            "kernel_ms": kernel_ms, "kernel_GBps": algo_bytes / (kernel_ms * 1e-3) / 1e9,
            "kernel_frac_of_hbm_peak": algo_bytes / (kernel_ms * 1e-3) / 1e9 / peaks["hbm"], "algorithmic_bytes": algo_bytes,
            "loader_ms_per_batch": loader_ms, "loader_sequences_per_s": B / (loader_ms * 1e-3),
            "h2d_bytes_per_batch": B * (16 * 8 + 4 + 10 * 9 * 4 + 10 * 8)})


def kernel_census(replay):
    """Informational: every kernel of ONE step (one graph replay) by name, from torch.profiler / CUPTI -- how many launches a
    step is, how much of its kernel time is this library's kernels and how much is torch glue (at::*).  Run after all the
    timed legs; never inside one."""
    import re
    import tempfile
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        replay()
        torch.cuda.synchronize()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "trace.json")
        prof.export_chrome_trace(path)
        events = json.load(open(path))["traceEvents"]
    ker = [e for e in events if e.get("cat") == "kernel" and "dur" in e]
    if os.environ.get("P2R_BENCH_CENSUS_DUMP") and ker:      # diagnostic: the step's kernels in time order, one per line
        k0 = min(e["ts"] for e in ker)
        with open(os.environ["P2R_BENCH_CENSUS_DUMP"], "w") as f:
            for e in sorted(ker, key=lambda e: e["ts"]):
                f.write("%9.1f %8.1f %4s %s\n" % (e["ts"] - k0, e["dur"], (e.get("args") or {}).get("stream"), e["name"][:110]))
    if not ker:
        return {"error": "no kernel events (CUPTI unavailable?)"}
    by = {}
    for e in ker:
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", e["name"])
        name = re.sub(r"^void ", "", name)
        key = name.split("(")[0][:72]
        c = by.setdefault(key, [0, 0.0])
        c[0] += 1
        c[1] += float(e["dur"])
    glue = [k for k in by if k.startswith("at::") or k.startswith("at_cuda") or "cub::" in k or k.startswith("nccl")]
    t0 = min(e["ts"] for e in ker)
    t1 = max(e["ts"] + e["dur"] for e in ker)
    streams = {}          # per stream: launches, busy time, first start and last end relative to the step's first kernel
    for e in ker:
        st = streams.setdefault(str((e.get("args") or {}).get("stream")), [0, 0.0, 1e30, 0.0])
        st[0] += 1
        st[1] += float(e["dur"])
        st[2] = min(st[2], e["ts"] - t0)
        st[3] = max(st[3], e["ts"] + e["dur"] - t0)
    # time with no kernel resident at all (launch / dependency bubbles) and with more than one (the overlap)
    edges = sorted([(e["ts"], 1) for e in ker] + [(e["ts"] + e["dur"], -1) for e in ker])
    depth, last, idle, multi = 0, t0, 0.0, 0.0
    for ts, d in edges:
        if depth == 0:
            idle += ts - last
        elif depth > 1:
            multi += ts - last
        depth += d
        last = ts
    top = sorted(by.items(), key=lambda kv: -kv[1][1])[:14]
    return {"kernels": len(ker), "kernel_time_us": sum(v[1] for v in by.values()), "span_us": t1 - t0,
            "torch_glue_kernels": sum(by[k][0] for k in glue), "torch_glue_time_us": sum(by[k][1] for k in glue),
            "memcpy_memset": sum(1 for e in events if e.get("cat") in ("gpu_memcpy", "gpu_memset")),
            "idle_us": round(idle, 1), "overlapped_us": round(multi, 1),
            "streams": [{"stream": k, "launches": v[0], "busy_us": round(v[1], 1), "first_us": round(v[2], 1),
                         "last_us": round(v[3], 1)} for k, v in sorted(streams.items(), key=lambda kv: -kv[1][1])],
            "top_by_time": [{"kernel": k, "launches": v[0], "us": round(v[1], 1)} for k, v in top],
            "glue_by_count": [{"kernel": k[:60], "launches": by[k][0], "us": round(by[k][1], 1)}
                              for k in sorted(glue, key=lambda k: -by[k][0])[:16]]}


# ------------------------------------------------------------------------------------------ our arm (GPU)
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.data_path_variants:
        return data_path_variants(args.batch)
    if args.leg:
        import bench_legs
        return bench_legs.main(args.leg)
    if int(os.environ.get("WORLD_SIZE", "1")) == 1 and os.environ.get("P2R_BENCH_CHILD") != "1" and \
            os.environ.get("P2R_BENCH_SUPERVISE", "1") != "0":
        sys.exit(supervise())
    start_watchdog()
    if os.environ.get("P2R_BENCH_TRACE_AFTER_S"):     # diagnostic: where is the host when a step does not come back
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["P2R_BENCH_TRACE_AFTER_S"]), repeat=False, file=sys.stderr)

    import torch.distributed as dist
    from pose2room_b200 import _lib, ops, synthetic
    from pose2room_b200.config import P2RConfig
    from pose2room_b200.p2rnet import P2RNet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    _lib.load()
    beat("library loaded")

    precision = args.precision
    if precision == "auto":
        try:
            from pose2room_b200 import gemm_sm100
            precision = "bf16" if gemm_sm100.available() else "fp32"
        except Exception:
            precision = "fp32"
    if precision == "bf16":
        from pose2room_b200 import gemm_sm100
        gemm_sm100.install()

    torch.manual_seed(42)
    np.random.seed(42)
    cfg = P2RConfig(mode="train", joint_num=JOINTS, num_frames=T_FRAMES, precision=precision)
    net = P2RNet(cfg)
    net.load_state_dict(synthetic.deterministic_state_dict(net.state_dict(), seed=7))
    net = net.to(dev).train()
    from pose2room_b200 import parallel
    parallel.broadcast_parameters(net)  # identical replicas, like DDP's initial broadcast
    params = [p for p in net.parameters() if p.requires_grad]
    # AdamW: the update rule of torch.optim.AdamW (what the reference's factory builds, models/optimizers.py:90) through
    # pose2room_b200.optim.AdamW -- a few launches, bias corrections once per thread (tests/test_optim_gpu.py holds it to
    # torch's results); P2R_FUSED_ADAMW=0 selects torch's capturable fused implementation (0.16 ms of the step)
    own_adamw = os.environ.get("P2R_FUSED_ADAMW", "1") != "0"
    if own_adamw:
        from pose2room_b200.optim import AdamW as _AdamW
        opt = _AdamW(params, lr=1e-3)
    else:
        opt = torch.optim.AdamW(params, lr=1e-3, fused=True, capturable=True)
    if precision == "bf16" and os.environ.get("P2R_WEIGHT_SHADOWS", "1") != "0":
        ops.register_weight_shadows(net)      # one multi-tensor fp32 -> bf16 copy per step instead of ~45 casts

    beat("model built")
    B = args.batch
    host = synthetic.make_batch(B, T_FRAMES, JOINTS, seed=1234 + rank, pin=True)
    tensors = {k: v for k, v in host.items() if isinstance(v, torch.Tensor)}
    resident = {k: v.to(dev) for k, v in tensors.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in tensors.values())

    overlap = os.environ.get("P2R_OVERLAP_DW", "1") != "0"
    dbg = (lambda *a: print("[bench rank %d]" % rank, *a, file=sys.stderr, flush=True)) if os.environ.get("P2R_BENCH_DEBUG") else (lambda *a: None)

    def fwd_bwd(data):
        opt.zero_grad(set_to_none=True)
        if overlap:
            with ops.overlap_weight_grads():   # dW GEMMs on a side stream, joined before the optimiser
                ep = net(data)
                loss = net.loss(ep, data)["total"]
                loss.backward()
        else:
            ep = net(data)
            loss = net.loss(ep, data)["total"]
            loss.backward()
        return loss

    def finish():
        """ONE flat gradient all-reduce over NCCL (the reference's DDP axis; no-op on one GPU) + AdamW."""
        parallel.allreduce_gradients(params)
        opt.step()

    def step(data):
        loss = fwd_bwd(data)
        finish()
        return loss

    # The whole step is ONE captured graph on every rank count: forward + loss + backward + the flat NCCL gradient
    # all-reduce + fused AdamW (NCCL collectives are capturable; the capture is thread-local so the process group's
    # watchdog thread cannot invalidate it).  Should that capture fail, the round-1 arrangement is the fallback: the graph
    # holds forward + loss + backward, the all-reduce runs eagerly after the replay and AdamW is a second small graph.
    graph_allreduce = world > 1 and os.environ.get("P2R_GRAPH_ALLREDUCE", "1") != "0"
    captured = step if (world == 1 or graph_allreduce) else fwd_bwd

    # ---- whole-step CUDA graph: ~3000 launches per step would otherwise be bound by the Python launch path ------
    static = {k: torch.empty_like(v) for k, v in resident.items()}
    for k in static:
        static[k].copy_(resident[k])
    graph, static_loss, use_graph = None, None, os.environ.get("P2R_CUDA_GRAPH", "1") != "0"
    first_step_loss = None      # loss of the very first step (same weights, same batch in every process: a parity signal)
    opt_graph = None
    if use_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(3):
                    l = step(static)
                    if i == 0:
                        first_step_loss = l.detach().clone()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            first_step_loss = float(first_step_loss.item())
            beat("eager warm-up done")
            dbg("eager warm-up done, capturing")
            graph = torch.cuda.CUDAGraph()
            opt.zero_grad(set_to_none=True)
            if graph_allreduce:
                ok = 1.0
                try:
                    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                        static_loss = captured(static)
                    torch.cuda.synchronize()
                except Exception as e:
                    print("bench.py: capture with the NCCL all-reduce inside failed (%r); all-reduce after the replay instead" % (e,),
                          file=sys.stderr)
                    ok = 0.0
                t_ok = torch.tensor([ok], device=dev)
                dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)          # every rank takes the same branch
                if float(t_ok.item()) < 1.0:
                    graph_allreduce, captured = False, fwd_bwd
                    graph = torch.cuda.CUDAGraph()
                    opt.zero_grad(set_to_none=True)
            if not graph_allreduce:
                with torch.cuda.graph(graph):
                    static_loss = captured(static)
            torch.cuda.synchronize()
            beat("graph captured")
            dbg("captured")
            if world > 1 and not graph_allreduce and os.environ.get("P2R_OPT_GRAPH", "1") != "0":
                # several ranks: the NCCL all-reduce stays outside the captures, but the fused AdamW (whose host-side
                # launch path costs more than its kernels) becomes a second small graph replayed right after it
                try:
                    graph.replay()                      # gradients exist at their captured addresses
                    parallel.allreduce_gradients(params)
                    opt_graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(opt_graph):
                        opt.step()
                    torch.cuda.synchronize()
                    beat("optimizer graph captured")
                except Exception as e:
                    print("bench.py: optimizer graph capture failed, stepping eagerly: %r" % (e,), file=sys.stderr)
                    opt_graph = None
        except Exception as e:  # report, do not hide: the bench line says whether the graph was used
            print("bench.py: CUDA graph capture failed, running eagerly: %r" % (e,), file=sys.stderr)
            graph, use_graph = None, False
    eager_step = step

    def step(data):  # noqa: F811  (graph replay with the batch copied into the captured buffers)
        if graph is None:
            if data is not static:
                data = {k: v.to(dev, non_blocking=True) for k, v in data.items()}
            return eager_step(data)
        if data is not static:
            for k in static:
                static[k].copy_(data[k], non_blocking=True)
        graph.replay()
        if world > 1 and not graph_allreduce:
            if opt_graph is not None:
                parallel.allreduce_gradients(params)
                opt_graph.replay()
            else:
                finish()
        return static_loss
    launches_per_step = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if graph is None:
        l0 = _lib.LAUNCHES["count"]
        eager_step(resident)
        launches_per_step = _lib.LAUNCHES["count"] - l0
    else:  # count once, eagerly, what the captured step launches through the C ABI
        l0 = _lib.LAUNCHES["count"]
        eager_step(static)
        launches_per_step = _lib.LAUNCHES["count"] - l0
    for _ in range(max(args.warmup, 3)):
        step(static)
    barrier()

    # ---- device-resident throughput -------------------------------------------------------------------
    beat("warm-up replays done")
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.timed(True)
    e0.record()
    for _ in range(args.steps):
        step(static)
    e1.record()
    barrier()
    sampler.timed(False)
    beat("timed region done")
    launches = launches_per_step * args.steps
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = world * B * args.steps / (ms * 1e-3)

    # per-GEMM device times: the same steps run eagerly with CUDA events around every GEMM launch (events cannot
    # be read back from inside a replayed graph); the clock sampler keeps running over both regions
    ops.PROFILE["log"] = []
    ops.PROFILE["on"] = True
    for _ in range(min(args.steps, 3)):
        eager_step(static)
    torch.cuda.synchronize()
    ops.PROFILE["on"] = False
    beat("per-kernel timing done")

    # ---- roofline of the dominant kernel: the fused graph-convolution GEMM (forward) -------------------
    vj = JOINTS * 64
    # the CTA-pair launch itself ("pair_fwd": events around the one C-ABI call); older paths log the whole operator ("fwd")
    gcn = [r for r in ops.PROFILE["log"] if r[0] == "pair_fwd" and r[2] == vj and r[3] == vj] or \
          [r for r in ops.PROFILE["log"] if r[0] == "fwd" and r[2] == vj and r[3] == vj]
    peaks = measured_peaks()
    roofline = None
    def live_fraction(kind):
        """Executed / dense FLOPs of the block-sparse graph-conv GEMMs (64x64 blocks of W_eff that are structurally zero
        are skipped): forward = the k-block lists of the 256-wide n-tiles, dW = the non-zero 128x128 output tiles."""
        try:
            from pose2room_b200 import gemm_sm100
            sp = net.backbone._w_sparsity
            if precision != "bf16" or not gemm_sm100.USE_SPARSITY:
                return 1.0
            if kind == "fwd":
                tab = sp.kb_list(gemm_sm100.PAIR_BLOCK_N, False, "cpu").numpy()
                widths = [min(gemm_sm100.PAIR_BLOCK_N, vj - i * gemm_sm100.PAIR_BLOCK_N) for i in range(tab.shape[0])]
                return float(sum(w * c * 64 for w, c in zip(widths, tab[:, 0]))) / float(vj * vj)
            bm, bn = DW_TILE
            mask = sp.tile_mask(bm, bn, "cpu").numpy()
            return float(mask.sum() * bm * bn) / float(vj * vj)      # (edge tiles counted whole: what the kernel executes)
        except Exception:
            return 1.0
    from pose2room_b200 import gemm_sm100 as _g
    DW_TILE = getattr(_g, "DW_TILE", (128, 128))
    if gcn:
        t_ms = sum(r[4].elapsed_time(r[5]) for r in gcn) / len(gcn)
        algo_flops = GCN_ALGO_GFLOP_PER_SEQ * 1e9 * B
        exec_flops = 2.0 * gcn[0][1] * vj * vj * live_fraction("fwd")
        achieved = algo_flops / (t_ms * 1e-3) / 1e12
        peak = peaks["tf_sustained"]
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": GCN_FWD_DRAM_BYTES_NCU if (precision == "bf16" and B == B_PER_GPU) else None,
                    "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r02_ncu_hot_kernels_summary.txt)",
                    "kernel": "graph-conv GEMM fwd (M=%d,N=K=%d) %s, CTA-pair tcgen05, block-sparse K, fused BN statistics" % (gcn[0][1], vj, precision),
                    "avg_launch_ms": t_ms, "launches_timed": len(gcn), "executed_tflops": exec_flops / (t_ms * 1e-3) / 1e12,
                    "executed_over_dense": live_fraction("fwd"),
                    "peak_source": peaks["src"] + " bf16 dense, sustained (kernel timed inside a long step)",
                    "share_of_step": t_ms * 6 / (ms / args.steps)}
    # the other large GEMM of a block: the graph convolution's weight gradient dW_eff = dG^T . X (same algorithmic FLOPs).
    # Timed here alone on the main stream; in the step it runs on the weight-gradient stream beside the BatchNorm chain.
    gdw = [r for r in ops.PROFILE["log"] if r[0] == "gcn_dw"]
    if gdw and roofline is not None:
        t_dw = sum(r[4].elapsed_time(r[5]) for r in gdw) / len(gdw)
        dw = {"bound": "tensor", "achieved": GCN_ALGO_GFLOP_PER_SEQ * 1e9 * B / (t_dw * 1e-3) / 1e12, "peak": peaks["tf_sustained"],
              "unit": "TFLOP/s", "kernel": "graph-conv weight gradient dW_eff[%d,%d] = dG^T.X over %d rows, %s" % (vj, vj, gdw[0][1], precision),
              "avg_launch_ms": t_dw, "launches_timed": len(gdw), "executed_over_dense": live_fraction("dw"),
              "executed_tflops": 2.0 * gdw[0][1] * vj * vj * live_fraction("dw") / (t_dw * 1e-3) / 1e12,
              "share_of_step": t_dw * 6 / (ms / args.steps),
              "traffic": None, "l2_to_sm_bytes": None}
        if precision == "bf16" and B == B_PER_GPU:     # ncu --set full of the same launch, profiles/r02_ncu_hot_kernels_summary.txt
            if getattr(_g, "USE_PAIR_DW", False):
                dw.update({"traffic": GCN_DW_PAIR_DRAM_BYTES_NCU, "l2_to_sm_bytes": 1.376e9,
                           "note": "CTA-pair kernel, 256 x 256 tiles (41 of 49 live), reduce-add split-K: tensor pipe 85.7 % active under ncu; "
                                   "1.38 GB through L2 -> SM per launch for 0.245 GB of DRAM traffic"})
            else:
                dw.update({"traffic": GCN_DW_DRAM_BYTES_NCU, "l2_to_sm_bytes": 1.83e9,
                           "note": "L2 -> SM bound: l1tex__m_xbar2l1tex_read_bytes 1.83 GB per launch for 0.21 GB of DRAM traffic (128 x 128 tiles)"})
        dw["frac"] = dw["achieved"] / dw["peak"]
        roofline["other_kernels"] = [dw]
    ops.PROFILE["log"] = []

    # ---- end to end: pinned host batch -> H2D -> step -> loss read back --------------------------------
    # Every step's batch is copied from pinned host memory inside the timed region and every step's loss is read back
    # (.item()).  Like a prefetching input pipeline (the reference's DataLoader uses pin_memory + workers,
    # dataloader.py:190-194), the H2D copy of batch i+1 runs on a copy stream while step i computes; the step itself
    # starts with a device-to-device copy from the landed staging buffers into the captured graph's input buffers.
    copy_stream = torch.cuda.Stream()
    staging = [{k: torch.empty_like(v) for k, v in resident.items()} for _ in range(2)]
    landed = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            for k in staging[i % 2]:
                staging[i % 2][k].copy_(tensors[k], non_blocking=True)
            landed[i % 2].record(copy_stream)

    def e2e_loop(n):
        """Synchronous variant: the host reads step i's loss (.item()) before it enqueues step i+1."""
        prefetch(0)
        for i in range(n):
            if i + 1 < n:
                prefetch(i + 1)
            torch.cuda.current_stream().wait_event(landed[i % 2])
            step(staging[i % 2]).item()      # D2D into the step's inputs -> step -> loss D2H (host sync)

    # Pipelined variant: the same copies and the same per-step loss read, but the loss of step i goes to pinned host
    # memory with an asynchronous copy and the host picks it up while step i+1 is already queued -- the GPU never waits
    # for the host round trip (what an asynchronous logger does; the reference's trainer calls .item(), training.py:42).
    loss_host, done, stepped = [None, None], [torch.cuda.Event(), torch.cuda.Event()], [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_loop_pipelined(n):
        losses = []
        prefetch(0)
        for i in range(n):
            cur = torch.cuda.current_stream()
            if i + 1 < n:
                if i >= 1:
                    copy_stream.wait_event(stepped[(i + 1) % 2])     # step i-1 has consumed the buffers batch i+1 lands in
                prefetch(i + 1)
            cur.wait_event(landed[i % 2])
            l = step(staging[i % 2]).detach()
            stepped[i % 2].record(cur)
            if loss_host[i % 2] is None:
                loss_host[i % 2] = torch.empty((), dtype=l.dtype).pin_memory()
            loss_host[i % 2].copy_(l, non_blocking=True)
            done[i % 2].record(cur)
            if i >= 1:
                done[(i - 1) % 2].synchronize()
                losses.append(float(loss_host[(i - 1) % 2]))
        done[(n - 1) % 2].synchronize()
        losses.append(float(loss_host[(n - 1) % 2]))
        return losses

    def time_e2e(loop):
        loop(2)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.timed(True)
        g0.record()
        loop(args.steps)
        g1.record()
        barrier()
        sampler.timed(False)
        return max_over_ranks(g0.elapsed_time(g1))

    ms_e2e_sync = time_e2e(e2e_loop)
    ms_e2e, e2e_mode = ms_e2e_sync, "synchronous (.item() after every step)"
    if os.environ.get("P2R_E2E_PIPELINED", "1") != "0":
        ok = 1.0
        try:
            ms_pipe = time_e2e(e2e_loop_pipelined)
        except Exception as e:
            print("bench.py: pipelined end-to-end loop failed, keeping the synchronous one: %r" % (e,), file=sys.stderr)
            ok, ms_pipe = 0.0, 0.0
        if world > 1:       # every rank must take the same branch
            ok = -max_over_ranks(-ok)
        # a loop that does not really wait would beat the device-resident step time: distrust it
        if ok > 0 and ms_pipe >= 0.97 * ms:
            ms_e2e, e2e_mode = ms_pipe, "pipelined by one step (async D2H of every step's loss into pinned memory; " \
                                        "the host reads loss i while step i+1 is queued)"
    sampler.stop_flag = True
    e2e = world * B * args.steps / (ms_e2e * 1e-3)
    beat("end-to-end leg done")

    data_path = None
    if rank == 0 and world == 1 and os.environ.get("P2R_BENCH_DATA_PATH", "1") != "0":
        try:
            data_path = data_path_leg(dev, B, train_step=step, steps=args.steps)
        except Exception as e:      # informational leg: report, never lose the headline line over it
            data_path = {"error": repr(e)}
        beat("data-path leg done")

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            seqs, _ = time_cpu_reference(2, 2, 1)
            cpu = {"value": seqs, "unit": "sequences/s", "cores": cores, "kind": "port",
                   "sample": "B=2 sequences x 2 steps of the same train step (T=1024, J=25) on %d host threads "
                             "(oracle/model_ref.py + oracle/pointnet2_ref.c)" % cores}
        line = {
            "metric": "pose-sequences/sec fwd+bwd (B=32, T=1024, J=25)", "value": value, "unit": "sequences/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": "P2RNet train step fwd+loss+bwd+AdamW (+grad all-reduce), B=%d/GPU, T=1024, J=25, "
                                   "512 seeds, 128 proposals, 22 classes" % B,
                       "parallelism": "dp%d" % world, "precision": precision, "cuda_graph": graph is not None, "overlap_dw": overlap,
                       "allreduce_in_graph": bool(world > 1 and graph is not None and graph_allreduce),
                       "optimizer": "pose2room_b200.optim.AdamW (torch.optim.AdamW's update rule, lr 1e-3, weight decay 1e-2)" if own_adamw
                                    else "torch.optim.AdamW(fused=True, capturable=True), lr 1e-3, weight decay 1e-2",
                       "l2": "no flush needed: per-layer activations (105-420 MB) exceed the 126 MB L2"},
            "clocks": sampler.summary(), "gpu_launches": launches,
            "e2e": {"value": e2e, "unit": "sequences/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": int(static_loss.element_size()) if static_loss is not None else 8,
                    "ms_per_step": ms_e2e / args.steps, "loss_read": e2e_mode,
                    "synchronous_ms_per_step": ms_e2e_sync / args.steps},
            "roofline": roofline, "cpu_baseline": cpu, "data_path": data_path, "first_step_loss": first_step_loss, "census": None,
        }
        if world == 1 and graph is not None and os.environ.get("P2R_BENCH_CENSUS", "1") != "0":
            # The census drives CUPTI over a graph replay.  The complete line goes out FIRST: should the profiler take the
            # process down, the supervisor keeps this line (it reads the last one; it prints exactly one).
            print(json.dumps(line), flush=True)
            beat("kernel census")
            _HEARTBEAT["limit"] = 45.0          # the line is out: a profiler that hangs costs 45 s, not the full stall limit
            try:
                line["census"] = kernel_census(lambda: step(static))
            except Exception as e:      # informational leg: report, never lose the headline line over it
                line["census"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Teardown: NCCL will not destroy a communicator while a captured graph still holds its kernels (round 2: with the
        # all-reduce inside the graph, destroy_process_group() sat there until the watchdog fired and torchrun reported a
        # failed rank although the line had been printed).  So: release the graphs first, and bound the teardown -- the
        # measurement is complete and printed, a stuck teardown must not turn the run into a failure.
        sys.stdout.flush()
        beat("shutdown")
        _HEARTBEAT["limit"], _HEARTBEAT["exit_code"] = 20.0, 0
        torch.cuda.synchronize()
        dist.barrier()
        for gr in (graph, opt_graph):
            if gr is not None:
                gr.reset()
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
