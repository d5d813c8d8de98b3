"""The benchmark driver and the measurement tools under tools/ must at least parse (they only run on a GPU box)."""
import ast
import glob
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_and_tools_parse():
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")] + sorted(glob.glob(os.path.join(ROOT, "tools", "*.py")))
    assert len(files) >= 8
    for f in files:
        ast.parse(open(f).read(), filename=f)


def test_bench_cli_contract():
    """bench.py keeps the driver's flags (--gpus/--steps/--warmup/--impl) and defaults to one GPU."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    for flag in ('"--gpus"', '"--steps"', '"--warmup"', '"--impl"'):
        assert flag in src
    assert 'ap.add_argument("--gpus", type=int, default=1)' in src


def test_tool_scripts_parse():
    """the sweep / validation shell scripts under tools/ (they run on a GPU box) are at least syntactically valid"""
    scripts = sorted(glob.glob(os.path.join(ROOT, "tools", "*.sh")))
    assert len(scripts) >= 6
    for f in scripts:
        r = subprocess.run(["bash", "-n", f], capture_output=True, text=True)
        assert r.returncode == 0, (f, r.stderr)
