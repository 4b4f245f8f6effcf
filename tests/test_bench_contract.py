"""bench.py: an extra config (BASELINE configs #3-#5, measured after the headline in the same process) that raises or
hangs must never cost the headline line -- checked on the CPU with the measurement itself stubbed out."""
import json
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

STUB = textwrap.dedent("""
    import sys, time, types
    sys.path.insert(0, {root!r})
    sys.argv = ["bench.py", "--steps", "2", "--warmup", "3"]
    import torch
    torch.cuda.is_available = lambda: True
    torch.cuda.empty_cache = lambda: None
    import bench
    bench.dist_setup = lambda args: (0, 1, 0, "cpu")
    def fake(args, rank, world, dev):
        base = {{"metric": "m", "value": 1.0, "unit": "sequences/s", "ms_per_step": 1.0, "steps": args.steps, "warmup": args.warmup,
                "dtype": "bf16", "gpu_launches": 1,
                "config": {{"global_batch": 1, "seq_len": 128, "local_batch": 1, "accumulation_steps": 1, "kfac": False,
                           "mlm_head": "x", "grad_reduction": "single"}},
                "clocks": {{}}, "extra": {{"loss_mean": 1.0}}}}
        if getattr(args, "roberta", False):
            raise RuntimeError("boom in an extra config")
        if getattr(args, "kfac", False) and {hang}:
            time.sleep(60)
        return base
    bench.run_config = fake
    bench.main()
""")


def _run(hang: bool, env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    code = STUB.format(root=ROOT, hang="True" if hang else "False")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, env=env)
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, (res.stdout[-500:], res.stderr[-1500:])
    return json.loads(lines[0]), res.returncode


def test_a_failing_extra_config_is_recorded_and_the_rest_still_runs():
    out, rc = _run(hang=False)
    assert rc == 0 and out["value"] == 1.0
    cfgs = out["extra"]["configs"]
    assert cfgs["phase2"]["value"] == 1.0 and "boom" in cfgs["roberta_fp8"]["error"] and cfgs["kfac"]["value"] == 1.0


def test_a_hanging_extra_config_times_out_and_the_headline_line_is_still_printed():
    out, rc = _run(hang=True, env_extra={"B200_BENCH_EXTRA_TIMEOUT": "2"})
    assert rc == 0 and out["value"] == 1.0
    cfgs = out["extra"]["configs"]
    assert cfgs["phase2"]["value"] == 1.0 and "timeout" in cfgs["kfac"]["error"]
