"""bench.py contract checks that need no GPU: the reference arm (the CPU restatement of the reference's path, the one
place besides tests/ and smoke() that may execute oracle/) prints ONE JSON line with the keys the driver reads, and
the B200 arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=300,
                          cwd=ROOT, env=env)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run(["--impl", "reference", "--reference-port", "--workload", "C1", "--obs-per-member", "2", "--steps", "1",
                "--warmup", "0", "--cpu-sample-members", "4", "--table-size", "200000"])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "perturbed-policy env-steps/sec" and d["unit"] == "env-steps/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env_rank = dict(RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=dict(os.environ, **env_rank))
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return          # on a GPU box the arm runs; the refusal is what is checked here
    out = _run(["--steps", "1", "--warmup", "1"])
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)


def test_precision_flag_maps_to_policy_levels():
    """--precision -> dfd_policy_desc.precision: the tensor paths are opt-out for IMPALA, opt-in below 32 observations per
    member for the MuJoCo MLPs, and never taken by the Discrete / Atari kernels (exact fp32 only)."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.forward_precision("mujoco", "auto", 128) == (True, 2)
    assert bench.forward_precision("mujoco", "auto", 1) == (False, 2)
    assert bench.forward_precision("mujoco", "tf32", 1) == (True, 1)
    assert bench.forward_precision("mujoco", "fp32", 128)[0] is False
    assert bench.forward_precision("impala", "auto", 1) == (True, 3)
    assert bench.forward_precision("impala", "fp32", 1) == (False, 3)
    assert bench.forward_precision("atari", "auto", 1) == (True, 1)
    assert bench.forward_precision("atari", "fp32", 1) == (False, 1)
    for prec in ("auto", "fp32", "tf32", "tf32a"):
        assert bench.forward_precision("discrete", prec, 128) == (False, 0)


def test_reference_arm_runs_the_reference_itself_when_it_is_importable():
    """with an importable copy of the unmodified reference (baseline/_ref, or $DFD_REFERENCE) the arm times
    the reference's own SharedNoiseTable / policy / FiniteDifferences.step / DSGD (`kind: "reference"`) and prints the
    number of steps it actually ran; the oracle port is the fallback only."""
    sys.path.insert(0, ROOT)
    import bench
    ref = bench.reference_root() or ("/root/reference" if os.path.exists("/root/reference/learner/finite_differences.py") else None)
    if ref is None:
        import pytest
        pytest.skip("no importable reference copy on this machine")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env["DFD_REFERENCE"] = ref
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1",
                          "--obs-per-member", "2", "--steps", "2", "--warmup", "1", "--cpu-sample-members", "4",
                          "--table-size", "200000"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["kind"] == "reference" and d["steps_run"] == 2 and d["value"] > 0
    assert "unmodified reference" in d["cpu_baseline"]["sample"]


def test_bench_parity_closed_form_equals_the_oracle():
    """bench.py's untimed parity step evaluates the estimator's fp64 closed form itself (the GPU arm never calls the
    oracle); here that evaluation is pinned to oracle.fd_gradient_closed_form on a small sharded batch - fd_return pairs
    and fd_state (delayed epochs) - by handing it a fake learner whose 'device' gradient IS the oracle's."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from oracle import dfd_oracle as O
    P, R, Hh = 777, 12, 4
    tab = O.NoiseTableOracle(50_000, P, 9)
    rng = np.random.RandomState(2)

    class T(object):
        _table = tab.table

    class Ctx(object):
        device = torch.device("cpu")

    class E(object):
        rank, world, pg, ctx = 0, 1, None, Ctx()
    for fd_state in (False, True):
        idx = rng.randint(0, 50_000 - P, size=R).astype(np.int64)
        rew = rng.randn(2 * R)
        sign = np.concatenate([np.ones(R), -np.ones(R)]).astype(np.int8)
        drows = (0.01 * rng.randn(Hh, P)).astype(np.float32)
        hrow = rng.randint(-1, Hh, size=2 * R).astype(np.int32)

        class L(object):
            pass
        L.P = P
        L.dist = torch.from_numpy(np.pad(drows, ((0, 0), (0, 3))))
        ref = O.fd_gradient_closed_form(tab.table, np.concatenate([idx, idx]), sign, rew, bench.SIGMA, P, 0.0,
                                        drows if fd_state else None, hrow if fd_state else None)

        def step():
            L.grad = torch.from_numpy(ref.astype(np.float32))
        out = bench.parity_step(E(), {"pairs": R}, T(), L, 0, [idx], torch.from_numpy(hrow) if fd_state else None,
                                torch.from_numpy(rew), step)
        assert out["ok"] and out["grad_rel_max"] < 2e-7 and out["ranks_bit_identical"] and out["coordinates_checked"] == P, out
        L.grad = torch.from_numpy((ref * 1.001).astype(np.float32))      # and it does notice a wrong gradient
        bad = bench.parity_step(E(), {"pairs": R}, T(), L, 0, [idx], torch.from_numpy(hrow) if fd_state else None,
                                torch.from_numpy(rew), lambda: None)
        assert not bad["ok"]
