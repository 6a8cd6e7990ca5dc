"""bench.py contract, CPU side: the reference arm (`--impl reference`) runs without a GPU, prints ONE JSON line
with the keys the driver reads, and under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

from conftest import ROOT

KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def run(cmd, env=None):
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    return [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_json_line():
    lines = run([sys.executable, "bench.py", "--impl", "reference", "--workload", "small", "--steps", "2", "--warmup", "1"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["metric"] == "lanczos_iterations_per_s" and d["unit"] == "iterations/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_only_rank0_prints():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29650")
    lines = run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--workload", "small", "--steps", "1", "--warmup", "0"], env)
    assert lines == []


def test_bench_functions_import_what_they_use():
    """bench.py imports numpy / torch lazily inside its functions (the reference arm must not need torch): every top-level
    function that names `np` or `torch` has to import it itself -- a leg that only runs on the GPU box would otherwise
    fail there with a NameError (this happened to the cfg5 leg)."""
    import ast
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    module_imports = {a.asname or a.name for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom)) for a in n.names}
    for f in [n for n in tree.body if isinstance(n, ast.FunctionDef)]:
        names = {n.id for n in ast.walk(f) if isinstance(n, ast.Name)}
        imps = {a.asname or a.name for n in ast.walk(f) if isinstance(n, (ast.Import, ast.ImportFrom)) for a in n.names}
        for alias in ("np", "torch"):
            assert alias not in names or alias in imps or alias in module_imports, "%s uses %s without importing it" % (f.name, alias)
