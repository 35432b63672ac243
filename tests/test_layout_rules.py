"""Repository rules that keep parity claims honest: the product never touches the oracle, the CUDA
extension is mandatory, and the CPU legs of bench.py are the only bench code that does."""
import ast
import importlib
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _imports(path: Path):
    tree = ast.parse(path.read_text())
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                yield a.name, node.lineno
        elif isinstance(node, ast.ImportFrom):
            yield (node.module or ""), node.lineno


def test_product_never_imports_the_oracle():
    for path in (ROOT / "quoracle_b200").rglob("*.py"):
        for mod, line in _imports(path):
            assert not mod.split(".")[0] == "oracle", f"{path}:{line} imports {mod}"


def test_csrc_does_not_reference_the_oracle():
    for path in (ROOT / "quoracle_b200" / "csrc").glob("*"):
        text = path.read_text()
        assert "qc_oracle" not in text
        assert not re.search(r"(?<!qu)oracle/", text), path      # "quoracle/" in reference citations is fine


def test_bench_touches_the_oracle_only_in_its_cpu_legs():
    tree = ast.parse((ROOT / "bench.py").read_text())
    allowed = {"cpu_path", "cpu_sim_only", "cpu_threads", "cpu_margins"}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef):
            for sub in ast.walk(node):
                if isinstance(sub, ast.ImportFrom) and (sub.module or "").split(".")[0] == "oracle":
                    assert node.name in allowed, f"bench.py:{sub.lineno} imports oracle in {node.name}()"
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    for n in top:
        mod = n.module if isinstance(n, ast.ImportFrom) else n.names[0].name
        assert (mod or "").split(".")[0] != "oracle"


def test_missing_extension_fails_loudly(monkeypatch, tmp_path):
    from quoracle_b200 import native
    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setattr(native, "_LIB_PATH", tmp_path / "libqconsensus.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        native.load_library()
    with pytest.raises(ImportError):
        native.Context(0)
    importlib.reload(native)


def test_oracle_header_says_test_infrastructure():
    assert "TEST INFRASTRUCTURE" in (ROOT / "oracle" / "__init__.py").read_text()
    assert "TEST INFRASTRUCTURE" in (ROOT / "oracle" / "qc_oracle.c").read_text()
