"""Static checks of bench.py's multi-rank control flow (no GPU needed).

One process per GPU: every collective (dist.barrier / all_reduce, and the helpers that wrap them) must be reached by
every rank.  A collective inside an `if rank == 0:` block deadlocks the job - which is what the default multi-GPU
command did while only `--no-extra` had been run on more than one GPU."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _src():
    with open(os.path.join(ROOT, "bench.py")) as f:
        return f.read()


def _is_rank0_test(node):
    return (isinstance(node, ast.Compare) and isinstance(node.left, ast.Name) and node.left.id == "rank"
            and len(node.ops) == 1 and isinstance(node.ops[0], ast.Eq)
            and isinstance(node.comparators[0], ast.Constant) and node.comparators[0].value == 0)


def _collective_calls(body):
    bad = []
    for stmt in body:
        for n in ast.walk(stmt):
            if not isinstance(n, ast.Call):
                continue
            f = n.func
            if isinstance(f, ast.Name) and f.id == "barrier":
                bad.append((n.lineno, "barrier()"))
            if isinstance(f, ast.Attribute) and isinstance(f.value, ast.Name) and f.value.id == "dist" \
                    and f.attr in ("barrier", "all_reduce", "all_gather", "broadcast", "reduce"):
                bad.append((n.lineno, f"dist.{f.attr}"))
            if isinstance(f, ast.Name) and f.id in ("max_over_ranks", "init_device_allreduce"):
                bad.append((n.lineno, f"{f.id}()"))
            if isinstance(f, ast.Name) and f.id == "timed":
                kw = {k.arg: k.value for k in n.keywords}
                v = kw.get("sync_all")
                if v is not None and not (isinstance(v, ast.Constant) and v.value is False):
                    bad.append((n.lineno, "timed(..., sync_all=True) calls barrier()"))
    return bad


def test_no_collective_inside_rank0_only_blocks():
    tree = ast.parse(_src())
    found, bad = 0, []
    for n in ast.walk(tree):
        if isinstance(n, ast.If) and _is_rank0_test(n.test):
            found += 1
            bad += _collective_calls(n.body)
    assert found >= 1
    assert not bad, f"collectives only rank 0 would reach: {bad}"


def test_bench_does_not_import_the_test_package():
    """bench.py is product measurement: its inputs come from adrates_b200.market_data, never from tests/."""
    tree = ast.parse(_src())
    for n in ast.walk(tree):
        if isinstance(n, ast.ImportFrom):
            assert not (n.module or "").startswith("tests"), n.lineno
        if isinstance(n, ast.Import):
            assert not any(a.name.startswith("tests") for a in n.names), n.lineno


def test_contract_keys_and_flags_are_present():
    src = _src()
    for key in ('"metric"', '"value"', '"unit"', '"n_gpus"', '"steps"', '"warmup"', '"ms_per_step"', '"higher_is_better"',
                '"scaling"', '"vs_baseline"', '"dtype"', '"data"', '"config"', '"clocks"', '"e2e"', '"gpu_launches"',
                '"roofline"', '"cpu_baseline"', '"impl"'):
        assert key in src, key
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in src, flag


def test_deadline_watchdog_ends_a_stalled_run():
    """BENCH_DEADLINE_S: a run that produces no result in time exits with 124 instead of hanging its launcher (checked
    on the CPU arm with an absurd step count)."""
    import subprocess
    import sys
    import time
    env = dict(os.environ, BENCH_DEADLINE_S="2")
    t0 = time.time()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "100000000",
                        "--warmup", "0"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=120)
    assert r.returncode == 124 and time.time() - t0 < 60
    assert "BENCH_DEADLINE_S" in r.stderr
