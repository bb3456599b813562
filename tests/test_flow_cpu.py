"""CPU: pins tests/commet_flow.py (the stand-in for Commet.py used on the GPU box) against the golden
outputs of the real Commet.py, both driving the reference's binaries (oracle/_ref)."""
import hashlib
import json
from pathlib import Path

import pytest

from oracle import oracle
from tests import commet_flow
from tests.golden import fixtures

GOLDEN = json.loads((Path(__file__).parent / "golden" / "golden.json").read_text())
pytestmark = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")


def check(work, res, g):
    for n in ("plain", "percentage", "normalized"):
        assert res[n] == g["csv"][n]
    got = {p.name: hashlib.sha256(p.read_bytes()).hexdigest() for p in (work / "output_commet").glob("*.bv")}
    assert got == g["bv"]


def test_flow_filtered_multichunk(tmp_path):
    fixtures.materialize(tmp_path)
    res = commet_flow.run("ABCDE_bench/sets_config.txt", oracle.REF_DIR, tmp_path, k=21, t=3, l=100, e=1.9, n=0, m=9000)
    check(tmp_path, res, GOLDEN["abcde_3sets_k21_filtered"])


def test_flow_abcde_k32(tmp_path):
    fixtures.materialize(tmp_path)
    res = commet_flow.run("ABCDE_bench/sets_config.txt", oracle.REF_DIR, tmp_path, k=32)
    check(tmp_path, res, GOLDEN["abcde_3sets_k32"])
