"""CPU: the product never routes through the test oracle, and fails loudly without its CUDA library.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/ (it is test infrastructure)."""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_nothing_under_the_package_references_the_oracle():
    offenders = []
    for p in (ROOT / "commet_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".hpp", ".h") and p.is_file():
            text = p.read_text(errors="replace")
            # imports, links, includes, paths or calls -- prose in comments may name the oracle
            if re.search(r"^\s*(from|import)\s+oracle\b|liboracle|commet_oracle|#include[^\n]*oracle|oracle[/.]\w+\s*\(|[\"']oracle[/\"']",
                         text, flags=re.M):
                offenders.append(str(p.relative_to(ROOT)))
    assert not offenders, offenders
    header = (ROOT / "include" / "commet_b200.h").read_text()
    assert "oracle" not in header


def test_loading_without_the_library_raises(tmp_path):
    """a copy of the Python binding without lib/libcommet_b200.so must raise, not fall back"""
    pkg = tmp_path / "commet_b200"
    pkg.mkdir()
    for name in ("__init__.py", "api.py", "multi.py", "build.py"):
        (pkg / name).write_text((ROOT / "commet_b200" / name).read_text())
    code = ("import commet_b200\n"
            "try:\n    commet_b200.load_library()\nexcept Exception as e:\n    print(type(e).__name__, e)\nelse:\n    print('LOADED')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "CommetError" in r.stdout and "no CPU fallback" in r.stdout


def test_context_creation_without_a_gpu_fails_with_a_message():
    """on a box without a CUDA device commet_ctx_create returns an error (there is no CPU path to fall into)"""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("this box has a GPU")
    import commet_b200
    try:
        commet_b200.Context(0)
    except commet_b200.CommetError as e:
        assert str(e)
    else:
        raise AssertionError("a context was created without a CUDA device")
