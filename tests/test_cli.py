"""GPLVMHaskell-exe twin (cli/gplvmhaskell_exe.cpp): Haskell `show` formatting (CPU) and the three-line CLI contract of
app/Main.hs:33-41 on the reference's fixture data (GPU)."""
import os
import re
import subprocess

import numpy as np
import pytest

from gplvmhaskell_b200 import build as gbuild
from oracle import ppca as oppca

HASKELL_SHOW = {  # GHC `show :: Double -> String`
    "0.1": "0.1", "1.0": "1.0", "5.0": "5.0", "-1.5": "-1.5", "0.05": "5.0e-2", "0.01": "1.0e-2", "123.456": "123.456",
    "12345000.0": "1.2345e7", "9999999.0": "9999999.0", "10000000.0": "1.0e7", "0.023676192353627": "2.3676192353627e-2",
    "-381.92805336": "-381.92805336", "5.1": "5.1", "0.003": "3.0e-3", "0.3333333333333333": "0.3333333333333333",
    "2147483562.0": "2.147483562e9", "1e22": "1.0e22", "5e-324": "5.0e-324",
}


def test_show_double_matches_haskell():
    exe = gbuild.build_cli()
    out = subprocess.run([exe, "--selftest-format"], capture_output=True, text=True, check=True).stdout.split()
    assert out == list(HASKELL_SHOW.values())
    for src, shown in HASKELL_SHOW.items():   # and every shown string reads back to the same double
        assert float(shown.replace("e", "E")) == float(src)


def test_cli_argument_errors():
    exe = gbuild.build_cli()
    assert subprocess.run([exe], capture_output=True).returncode != 0
    assert subprocess.run([exe, "nofile.txt", "10", "Maybe"], capture_output=True).returncode != 0


def _write_fixture(path, Y, nan_token="NaN", trailing_newline=True):
    lines = []
    for n in range(Y.shape[1]):
        lines.append("  ".join(nan_token if np.isnan(v) else repr(float(v)) for v in Y[:, n]))
    txt = "\n".join(lines) + ("\n" if trailing_newline else "")
    open(path, "w").write(txt)


def _parse_show_array(line):
    m = re.match(r"^(Just \()?AUnboxed \(\(Z :\. (\d+)\) :\. (\d+)\) \[(.*)\]\)?$", line)
    assert m, line[:120]
    vals = np.array([float(v) for v in m.group(4).split(",")])
    return vals.reshape(int(m.group(2)), int(m.group(3)))


@pytest.mark.gpu
@pytest.mark.parametrize("typesafe", ["False", "True"])
def test_cli_dense_config1(tmp_path, iris, typesafe):
    """GPLVMHaskell-exe test.pca.txt 1000 False: three stdout lines, line 3 is Nothing."""
    exe = gbuild.build_cli()
    f = tmp_path / "test.pca.txt"
    _write_fixture(f, iris["Y"])
    r = subprocess.run([exe, str(f), "1000", typesafe, "--seed", "7", "--sigma2", "1.0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().split("\n")
    assert len(lines) == 3 and lines[2] == "Nothing"
    W = _parse_show_array(lines[0])
    assert W.shape == (4, 3)
    # the fixed point does not depend on W0: compare with the oracle's 1001-step run
    assert abs(float(lines[1]) - float(iris["dense_s1_k1000_ll"])) < 1e-6 * abs(float(iris["dense_s1_k1000_ll"]))
    assert np.max(np.abs(oppca.projector(W) - oppca.projector(iris["dense_s1_k1000_W"]))) < 1e-4


@pytest.mark.gpu
def test_cli_missing_config2(tmp_path, iris):
    """GPLVMHaskell-exe test.pca.nan.txt 1000 True: line 3 is Just (AUnboxed ((Z :. 4) :. 150) [...])."""
    exe = gbuild.build_cli()
    f = tmp_path / "test.pca.nan.txt"
    _write_fixture(f, iris["Yn"], trailing_newline=False)   # the reference's NaN fixture has no trailing newline
    r = subprocess.run([exe, str(f), "1000", "True", "--seed", "7", "--sigma2", "1.0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().split("\n")
    assert len(lines) == 3 and lines[2].startswith("Just (AUnboxed ((Z :. 4) :. 150) [")
    R = _parse_show_array(lines[2])
    assert R.shape == (4, 150) and np.isfinite(R).all()
    assert abs(float(lines[1]) - float(iris["miss_s1_k1000_ll"])) < 1e-5 * abs(float(iris["miss_s1_k1000_ll"]))
    obs = ~np.isnan(iris["Yn"])
    assert np.sqrt(np.mean((R[obs] - iris["Yn"][obs]) ** 2)) < 0.2
