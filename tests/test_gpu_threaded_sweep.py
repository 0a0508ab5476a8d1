"""parallel.threaded_sweep: the q values of a sweep side by side on one GPU (a host thread and a stream per q) give what
the same runs give one after another."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_threaded_q_sweep_equals_sequential(gpu_required):
    from graphbix_b200.engine import EM, EM_mod
    from graphbix_b200.parallel import threaded_sweep
    from graphbix_b200.synth import planted_partition
    links, _ = planted_partition(3000, 3, 8.0, 0.1, seed=5, connected_only=True)
    n = int(links.max())
    qs = [2, 3, 4, 5, 6]
    run_mod = lambda q, st: EM_mod(np.ones(q) / q, n, q, links, 1, 200, 0.0, True, True, False, 0.3, seed=5 + q, stream=st)
    run_sbm = lambda q, st: EM(n, q, links, n, "uniformAssortative", False, 100, True, 0.3, seed=q, stream=st)
    for run, ipsi, iitr in ((run_mod, 5, 16), (run_sbm, 2, 14)):
        seq = [run(q, None) for q in qs]
        for rep in range(3):   # thread interleavings differ from run to run
            par = threaded_sweep(qs, run)
            for a, b in zip(seq, par):
                assert a[iitr] == b[iitr]
                np.testing.assert_array_equal(a[ipsi], b[ipsi])
