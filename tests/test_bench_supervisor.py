"""CPU: bench.py's supervisor -- a measuring child that stalls is killed and the SAME configuration is measured once more,
flagged in config.retry; a child that fails for another reason is not retried."""
import json
import os
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run(tmp_path, body, monkeypatch, capsys, timeout="3"):
    import bench
    script = tmp_path / "fake_child.py"
    script.write_text(textwrap.dedent(body))
    monkeypatch.setenv("P2R_BENCH_TIMEOUT_S", timeout)
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    rc = bench.supervise(str(script))
    return rc, capsys.readouterr()


def test_stalled_child_is_retried_once_with_the_same_configuration(tmp_path, monkeypatch, capsys):
    marker = tmp_path / "first_attempt_seen"
    monkeypatch.setenv("FAKE_MARKER", str(marker))
    rc, io = _run(tmp_path, '''
        import json, os, time
        assert os.environ["P2R_BENCH_CHILD"] == "1" and os.environ.get("P2R_OVERLAP_DW") is None
        if not os.path.exists(os.environ["FAKE_MARKER"]):
            open(os.environ["FAKE_MARKER"], "w").close()
            time.sleep(60)                       # the first attempt never finishes
        print(json.dumps({"value": 1.0, "config": {"workload": "w"}}))
    ''', monkeypatch, capsys)
    assert rc == 0
    lines = [l for l in io.out.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                         # exactly ONE json line
    d = json.loads(lines[0])
    assert d["value"] == 1.0 and "same configuration" in d["config"]["retry"]
    assert "timed out" in io.err


def test_watchdog_exit_is_retried_and_real_failures_are_not(tmp_path, monkeypatch, capsys):
    marker = tmp_path / "watchdog_fired_once"
    monkeypatch.setenv("FAKE_MARKER", str(marker))
    rc, io = _run(tmp_path, '''
        import json, os, sys
        if not os.path.exists(os.environ["FAKE_MARKER"]):
            open(os.environ["FAKE_MARKER"], "w").close()
            sys.exit(17)                         # what the in-process watchdog does on a stall
        print(json.dumps({"value": 2.0, "config": {}}))
    ''', monkeypatch, capsys, timeout="30")
    assert rc == 0 and json.loads(io.out.strip().splitlines()[-1])["config"]["retry"]
    rc, io = _run(tmp_path, '''
        import sys
        print("bench.py: no CUDA device")
        sys.exit(1)
    ''', monkeypatch, capsys, timeout="30")
    assert rc == 1 and "no CUDA device" in io.out and not any(l.startswith("{") for l in io.out.splitlines())


def test_healthy_child_line_is_passed_through_unchanged(tmp_path, monkeypatch, capsys):
    rc, io = _run(tmp_path, '''
        print('{"value": 3.5, "config": {"workload": "w"}}')
    ''', monkeypatch, capsys, timeout="30")
    assert rc == 0 and io.out.strip() == '{"value": 3.5, "config": {"workload": "w"}}'


def test_variant_ab_runs_in_its_own_process_and_cannot_lose_the_line(tmp_path, monkeypatch, capsys):
    """The kernel-variant A/B of the sample -> batch path is merged into data_path.variants when it succeeds and becomes
    an {"error": ...} entry when its process crashes, prints nothing or hangs -- the headline line survives every time."""
    body = '''
        import json, os, sys, time
        if "--data-path-variants" in sys.argv:
            mode = os.environ["FAKE_VARIANTS"]
            if mode == "ok":
                print(json.dumps({"v1": {"kernel_ms": 0.05}, "v2": {"kernel_ms": 0.02, "bit_identical_to_v1": True}}))
            elif mode == "crash":
                os.abort()
            elif mode == "hang":
                time.sleep(60)
            sys.exit(0)
        print(json.dumps({"value": 4.0, "config": {}, "data_path": {"kernel_ms": 0.05}}))
    '''
    monkeypatch.setenv("FAKE_VARIANTS", "ok")
    rc, io = _run(tmp_path, body, monkeypatch, capsys, timeout="30")
    d = json.loads(io.out.strip())
    assert rc == 0 and d["value"] == 4.0 and d["data_path"]["variants"]["v2"]["bit_identical_to_v1"] is True
    monkeypatch.setenv("FAKE_VARIANTS", "crash")
    rc, io = _run(tmp_path, body, monkeypatch, capsys, timeout="30")
    d = json.loads(io.out.strip())
    assert rc == 0 and d["value"] == 4.0 and "error" in d["data_path"]["variants"]
    monkeypatch.setenv("FAKE_VARIANTS", "hang")
    monkeypatch.setenv("P2R_BENCH_VARIANTS_TIMEOUT_S", "2")
    rc, io = _run(tmp_path, body, monkeypatch, capsys, timeout="30")
    d = json.loads(io.out.strip())
    assert rc == 0 and d["value"] == 4.0 and d["data_path"]["variants"] == {"error": "timed out"}
    # a failed data-path leg in the child (or none at all) starts no second process
    rc, io = _run(tmp_path, '''
        import sys
        assert "--data-path-variants" not in sys.argv
        print('{"value": 5.0, "config": {}, "data_path": {"error": "boom"}}')
    ''', monkeypatch, capsys, timeout="30")
    assert rc == 0 and json.loads(io.out.strip())["data_path"] == {"error": "boom"}


def test_experiments_and_legs_report_next_to_the_headline_and_cannot_lose_it(tmp_path, monkeypatch, capsys):
    """Alternative configurations (fp32 parity mode, the unfused loss / head paths) and the secondary legs (the reference on
    the same GPU, forward only / SA operators / eval) are measured by fresh children after the headline is safe: results (or
    an error / a skip) land next to it, the headline numbers stay the first child's."""
    body = '''
        import json, os, sys, time
        if "--leg" in sys.argv:
            leg = sys.argv[sys.argv.index("--leg") + 1]
            if leg == "gpu_reference":
                print(json.dumps({"fp32": {"train_sequences_per_s": 2.0, "forward_ms": 7.0},
                                  "tf32": {"train_sequences_per_s": 4.0, "forward_ms": 6.0}}))
            elif os.environ.get("FAKE_EXTRAS") == "crash":
                os.abort()
            else:
                print(json.dumps({"forward_only": {"bf16": {"forward_ms": 1.0}}, "sa_operators": [], "sa_module_forward": {},
                                  "eval_1k": {"total_ms_per_scene": 0.5}}))
            sys.exit(0)
        if os.environ.get("P2R_BENCH_DATA_PATH") == "0":           # an experiment child
            assert "--no-cpu-baseline" in sys.argv and sys.argv[sys.argv.index("--steps") + 1] == "2"
            fp32 = "--precision" in sys.argv and sys.argv[sys.argv.index("--precision") + 1] == "fp32"
            unfused = all(os.environ.get(k) == "0" for k in ("P2R_FUSED_LOSS", "P2R_FUSED_GMM", "P2R_FUSED_VOTE"))
            assert fp32 != unfused
            if unfused and os.environ.get("FAKE_UNFUSED") == "crash":
                os.abort()
            print(json.dumps({"value": 1.0 if fp32 else 8.0, "ms_per_step": 64.0 if fp32 else 2.0, "first_step_loss": 0.5,
                              "dtype": "f32" if fp32 else "bf16", "gpu_launches": 7, "config": {"cuda_graph": True},
                              "e2e": {"ms_per_step": 3.0}}))
            sys.exit(0)
        print(json.dumps({"value": 4.0, "ms_per_step": 5.0, "steps": 2, "warmup": 3, "n_gpus": 1, "config": {},
                          "roofline": {"frac": 0.5}, "first_step_loss": 0.5}))
    '''
    rc, io = _run(tmp_path, body, monkeypatch, capsys, timeout="30")
    d = json.loads(io.out.strip())
    ex = d["experiments"]
    assert rc == 0 and d["value"] == 4.0 and ex["baseline"] == {"ms_per_step": 5.0, "first_step_loss": 0.5, "kernels_per_step": None}
    assert ex["fp32_mode"]["ms_per_step"] == 64.0 and ex["fp32_mode"]["dtype"] == "f32"
    assert ex["unfused_loss+gmm+vote"]["e2e_ms_per_step"] == 3.0
    assert d["gpu_reference"]["tf32"]["train_sequences_per_s"] == 4.0
    assert d["vs_gpu_reference"]["bf16_over_reference_fp32"] == 2.0 and d["vs_gpu_reference"]["fp32_mode_over_reference_fp32"] == 0.5
    assert d["forward_only"]["bf16"]["forward_ms"] == 1.0 and d["forward_only"]["gpu_reference_fp32_forward_ms"] == 7.0
    assert d["eval_1k"]["total_ms_per_scene"] == 0.5
    monkeypatch.setenv("FAKE_UNFUSED", "crash")
    monkeypatch.setenv("FAKE_EXTRAS", "crash")
    rc, io = _run(tmp_path, body, monkeypatch, capsys, timeout="30")
    d = json.loads(io.out.strip())
    assert rc == 0 and d["value"] == 4.0 and "error" in d["experiments"]["unfused_loss+gmm+vote"] and "forward_only" not in d
    monkeypatch.setenv("P2R_BENCH_EXPERIMENTS_BUDGET_S", "1")      # no time left: every experiment is skipped, line intact
    rc, io = _run(tmp_path, body, monkeypatch, capsys, timeout="30")
    d = json.loads(io.out.strip())
    assert rc == 0 and d["value"] == 4.0 and all("skipped" in v for k, v in d["experiments"].items() if k != "baseline")
    monkeypatch.setenv("P2R_BENCH_EXPERIMENTS", "0")
    monkeypatch.setenv("P2R_BENCH_LEGS", "0")
    rc, io = _run(tmp_path, body, monkeypatch, capsys, timeout="30")
    d = json.loads(io.out.strip())
    assert rc == 0 and d["experiments"] is None and d["gpu_reference"] is None


def test_child_that_dies_in_the_census_leg_keeps_its_line(tmp_path, monkeypatch, capsys):
    """The measuring child prints its complete line before it drives CUPTI; if the profiler takes the process down the
    supervisor still reports that line (once), with the census marked as failed."""
    rc, io = _run(tmp_path, '''
        import json, os, sys
        print(json.dumps({"value": 6.0, "config": {}, "census": None}), flush=True)
        os.abort()
    ''', monkeypatch, capsys, timeout="30")
    lines = [l for l in io.out.splitlines() if l.startswith("{")]
    d = json.loads(lines[0])
    assert rc == 0 and len(lines) == 1 and d["value"] == 6.0 and "error" in d["census"]
    rc, io = _run(tmp_path, '''
        import json
        print(json.dumps({"value": 6.0, "config": {}, "census": None}), flush=True)
        print(json.dumps({"value": 6.0, "config": {}, "census": {"kernels": 1234}}), flush=True)
    ''', monkeypatch, capsys, timeout="30")
    lines = [l for l in io.out.splitlines() if l.startswith("{")]
    assert rc == 0 and len(lines) == 1 and json.loads(lines[0])["census"] == {"kernels": 1234}


def test_child_that_hangs_in_the_census_leg_keeps_its_line(tmp_path, monkeypatch, capsys):
    rc, io = _run(tmp_path, '''
        import json, time
        print(json.dumps({"value": 7.0, "config": {}, "census": None}), flush=True)
        time.sleep(60)
    ''', monkeypatch, capsys, timeout="3")
    lines = [l for l in io.out.splitlines() if l.startswith("{")]
    d = json.loads(lines[0])
    assert rc == 0 and len(lines) == 1 and d["value"] == 7.0 and "error" in d["census"] and "retry" not in d["config"]
