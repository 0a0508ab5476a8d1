"""The C ABI used from plain C (no Python, no torch in the process): tests/cabi/cabi_client.c, compiled with gcc
against include/graphbix_cuda.h, must give what the ctypes path gives on the same problem."""
import os
import struct
import subprocess

import numpy as np
import pytest

from helpers import assortative_init, planted_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_program_matches_ctypes_path(gpu_required, tmp_path):
    from graphbix_b200 import _lib
    from graphbix_b200.engine import SbmEngine
    lib_path = _lib.LIB_PATH
    exe = tmp_path / "cabi_client"
    subprocess.run(["gcc", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cabi", "cabi_client.c"),
                    "-o", str(exe), "-L", os.path.dirname(lib_path), "-lgraphbix_cuda",
                    f"-Wl,-rpath,{os.path.dirname(lib_path)}"], check=True)
    q = 3
    n, links, psi0, _ = planted_case(900, q, 9.0, 0.1, seed=3)
    gr0, C0 = assortative_init(q)
    K = links.shape[0]
    src, dst = tmp_path / "problem.bin", tmp_path / "result.bin"
    with open(src, "wb") as f:
        f.write(struct.pack("<3q", n, K, q))
        f.write(np.asfortranarray(links, dtype=np.int64).tobytes(order="F"))
        f.write(np.asarray(gr0, dtype=np.float64).tobytes())
        f.write(np.asarray(C0, dtype=np.float64).tobytes())
        f.write(np.asarray(psi0, dtype=np.float64).tobytes())
    out = subprocess.run([str(exe), str(src), str(dst)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    raw = open(dst, "rb").read()
    scal = np.frombuffer(raw, dtype=np.float64, count=12)
    off = 12 * 8
    gr = np.frombuffer(raw, dtype=np.float64, count=q, offset=off); off += 8 * q
    Crs = np.frombuffer(raw, dtype=np.float64, count=q * q, offset=off).reshape(q, q); off += 8 * q * q
    PSI = np.frombuffer(raw, dtype=np.float64, count=n * q, offset=off).reshape(n, q); off += 8 * n * q
    block = np.frombuffer(raw, dtype=np.int32, count=n, offset=off)

    eng = SbmEngine(n, links, q)
    eng.set_options(itrmax=300)
    eng.init(gr0, C0, psi0)
    res = eng.em()
    st = eng.state()
    assert scal[5] == res.cnv == 1 and int(scal[6]) == res.itrnum
    np.testing.assert_allclose(scal[:5], [res.FE, res.CVBayes, res.CVGP, res.CVGT, res.CVMAP], rtol=1e-13)
    np.testing.assert_allclose(gr, st["gr"], rtol=1e-13)
    np.testing.assert_allclose(Crs, st["Crs"], rtol=1e-13)
    np.testing.assert_allclose(PSI, st["PSI"], rtol=0, atol=1e-14)
    np.testing.assert_array_equal(block, eng.assignment())
    assert scal[10] > 0                      # kernels were launched by the C process
    assert int(scal[11]) == -1               # GBX_ERR_INVALID for the self loop
