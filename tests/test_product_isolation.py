"""CPU: the oracle is test infrastructure.  Nothing under vitta_b200/ (the product), nor the entry scripts, may import it;
bench.py may only do so inside its CPU-baseline leg; tools/ that do are study tools, not on any product path."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _imports(path):
    tree = ast.parse(open(path).read())
    mods = []
    for n in ast.walk(tree):
        if isinstance(n, ast.Import):
            mods += [(a.name, n.lineno) for a in n.names]
        elif isinstance(n, ast.ImportFrom):
            mods.append(((n.module or "") if n.level == 0 else "." * n.level + (n.module or ""), n.lineno))
    return mods


def test_product_package_never_imports_the_oracle_or_the_reference():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vitta_b200")):
        for f in files:
            if f.endswith(".py"):
                for mod, line in _imports(os.path.join(dirpath, f)):
                    assert not mod.split(".")[0] in ("oracle", "tests", "cases"), (dirpath, f, line, mod)
    for script in ("tta_tanet_ucf101.py", "tta_swin_ucf101.py"):
        for mod, line in _imports(os.path.join(ROOT, script)):
            assert mod.split(".")[0] != "oracle", (script, line)


def test_bench_touches_the_oracle_only_in_the_cpu_baseline_leg():
    path = os.path.join(ROOT, "bench.py")
    tree = ast.parse(open(path).read())
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        uses = [n for n in ast.walk(fn) if isinstance(n, (ast.Import, ast.ImportFrom))
                and any("oracle" in (getattr(n, "module", None) or "") or "oracle" in a.name for a in n.names)]
        if uses:
            assert fn.name == "_oracle_tanet_state", fn.name
    # ... and that loader is only called by the baseline legs (CPU port, configs[0] CPU forward, reference-on-GPU record:
    # baselines measured BESIDE the product, never the product path)
    callers = {fn.name for fn in ast.walk(tree) if isinstance(fn, ast.FunctionDef)
               for n in ast.walk(fn) if isinstance(n, ast.Call) and getattr(n.func, "id", None) == "_oracle_tanet_state"}
    assert callers == {"cpu_port_clips_per_s", "cpu_cfg1_eval_clips_per_s", "gpu_reference_step"}, callers
    baseline_legs = callers | {"_oracle_tanet_state"}
    product = {"build_tanet", "build_swin", "secondary_records", "parity_check", "attribute_step", "stats_kernel_roofline"}
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name in product]:
        names = {getattr(n.func, "id", None) for n in ast.walk(fn) if isinstance(n, ast.Call)}
        assert not (names & baseline_legs), (fn.name, names & baseline_legs)
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    assert not any("oracle" in (getattr(n, "module", None) or "") for n in top)


def test_no_reference_path_at_run_time():
    """/root/reference does not exist on the GPU box: only the golden generators under oracle/ may name it."""
    for rel in ["bench.py", "__graft_entry__.py"] + [os.path.join("tests", f) for f in os.listdir(os.path.join(ROOT, "tests"))
                                                    if f.endswith(".py") and f != "test_product_isolation.py"]:
        assert "/root/reference" not in open(os.path.join(ROOT, rel)).read(), rel
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vitta_b200")):
        for f in files:
            if f.endswith(".py"):
                assert "/root/reference" not in open(os.path.join(dirpath, f)).read(), f
