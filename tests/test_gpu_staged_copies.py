"""Large PAGEABLE host arrays (what a Julia `ccall` hands over) cross PCIe through pinned chunks filled by several host
threads (gbx_pool.cu: upload_async / download_sync).  The bytes that arrive must be the bytes of the plain copies: the
edge list (create), the initial messages (init), PSI (get_state) and the messages (get_messages)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(n, links, psi0, q, threads, min_mb):
    from graphbix_b200 import _lib
    from graphbix_b200._lib import check
    from graphbix_b200.engine import SbmEngine
    from helpers import assortative_init
    os.environ["GBX_STAGE_THREADS"] = str(threads)
    os.environ["GBX_STAGE_MIN_MB"] = str(min_mb)
    try:
        gr0, C0 = assortative_init(q)
        e = SbmEngine(n, links, q)          # links: a numpy array, i.e. pageable
        e.init(gr0, C0, psi0)               # K x q doubles, pageable
        e.iterate(3)
        lib = _lib.load()
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        PSI = np.full((n, q), -1.0)         # plain numpy on purpose: engine.state() would hand out pinned arrays
        th = np.full(n, -1.0)
        gr, Crs, h = np.empty(q), np.empty((q, q)), np.empty(q)
        check(lib.gbx_sbm_get_state(e._h, dp(gr), dp(Crs), dp(PSI), dp(th), dp(h)))
        msg = e.messages()
        e.close()
        return PSI, th, Crs, msg
    finally:
        os.environ.pop("GBX_STAGE_THREADS", None)
        os.environ.pop("GBX_STAGE_MIN_MB", None)


@pytest.mark.parametrize("threads", [1, 3])
def test_staged_copies_move_the_same_bytes(gpu_required, threads):
    from helpers import planted_case
    # 1.6M links x 2 x 8 B = 26 MB = 13 chunks of 2 MB (ragged last one); messages 1.6M x 4 x 8 B = 51 MB; PSI 6.4 MB
    n, links, psi0, _ = planted_case(200_000, 4, 16.0, 0.1, seed=77)
    assert links.shape[0] * 16 > 6 * (4 << 20)
    plain = _run(n, links, psi0, 4, 0, 16)
    staged = _run(n, links, psi0, 4, threads, 0)
    for a, b in zip(plain, staged):
        assert (a != -1.0).all()
        np.testing.assert_array_equal(a, b)


def test_staged_upload_keeps_the_error_checks(gpu_required):
    """A bad edge list is still refused when it arrives through the staged path."""
    from graphbix_b200._lib import GbxError
    from graphbix_b200.engine import SbmEngine
    from helpers import planted_case
    n, links, _, _ = planted_case(50_000, 2, 12.0, 0.1, seed=5)
    bad = links.copy()
    bad[len(bad) // 2, 0] = n + 7   # out of range, in the middle of the list
    os.environ["GBX_STAGE_THREADS"] = "2"
    os.environ["GBX_STAGE_MIN_MB"] = "0"
    try:
        with pytest.raises(GbxError):
            SbmEngine(n, bad, 2)
        SbmEngine(n, links, 2).close()
    finally:
        os.environ.pop("GBX_STAGE_THREADS", None)
        os.environ.pop("GBX_STAGE_MIN_MB", None)
