"""CPU: host-side pieces of commet_nxn that need no GPU -- the Python-3 float formatting of the CSV matrices
(Commet.py:298,313 write str(float)) is compiled out of the tool's source and compared with Python itself."""
import random
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_py_float_matches_python_str(tmp_path):
    src = (ROOT / "commet_b200" / "csrc" / "tools" / "commet_nxn.cpp").read_text()
    fn = src[src.index("static std::string py_float(double x)"):src.index("static void ensure_dir")]
    (tmp_path / "t.cpp").write_text(
        "#include <charconv>\n#include <cmath>\n#include <string>\n#include <cstdio>\n#include <cstdlib>\n" + fn +
        'int main(){double x; while(scanf("%la",&x)==1) puts(py_float(x).c_str());}\n')
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", str(tmp_path / "t"), str(tmp_path / "t.cpp")], check=True)
    rnd = random.Random(1)
    vals = [0.0, 100.0, 25.0, 100 / 3, 1e-5, 5e-06, 1e16, 1e15, 123456789012345678.0, 0.0001, 0.00012345, 1.5e-7,
            200 / 3, 99.99999999999999, 1e22, 3.0e-310, 100 * 1 / float(20_000_000), 100 * 7 / float(500_000_000)]
    for _ in range(5000):
        c, n = rnd.randint(0, 10 ** rnd.randint(1, 9)), rnd.randint(1, 10 ** rnd.randint(1, 9))
        vals += [100 * c / float(n), 100 * (c + n) / float(n + rnd.randint(1, 10 ** 9))]
    out = subprocess.run([str(tmp_path / "t")], input="\n".join(float.hex(v) for v in vals), capture_output=True,
                         text=True, check=True).stdout.split("\n")
    bad = [(v, o) for v, o in zip(vals, out) if o != str(v)]
    assert not bad, bad[:5]
