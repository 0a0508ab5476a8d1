"""The partition C ABI (gbx_sbm_create_part / gbx_sbm_step / ...) and its driver graphbix_b200.engine_dist.PartEngine.

On a one-GPU box: a single-rank "partition" must reproduce SbmEngine / ModEngine step for step (same kernels, the
collectives are identities).  With two or more GPUs the two-rank check tests/dist_check_part.py runs under torchrun
(packed and direct boundary exchange, several pieces, hubs, q = 2..8, both models, against the single-GPU engine)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import assortative_init, planted_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def single_rank_group():
    import torch
    import torch.distributed as dist
    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        created = True
    yield dist
    if created:
        dist.destroy_process_group()


@pytest.mark.parametrize("q,dc", [(3, True), (8, False)])
def test_single_rank_partition_equals_engine(gpu_required, single_rank_group, q, dc):
    from graphbix_b200.engine import SbmEngine
    from graphbix_b200.engine_dist import PartEngine
    n, links, psi0, _ = planted_case(3000, q, 9.0, 0.1, seed=50 + q)
    gr0, C0 = assortative_init(q)
    ref = SbmEngine(n, links, q)
    ref.set_options(degreecorrection=int(dc))
    ref.init(gr0, C0, psi0)
    part = PartEngine(n, links, q, "sbm")
    assert (part.vbeg, part.n, part.sbeg, part.K) == (0, n, 0, links.shape[0])
    part.set_options(degreecorrection=int(dc))
    part.init(gr0, C0, psi0)
    for _ in range(5):
        ref.iterate(1)
        part.iterate(1)
        sr, sp = ref.state(), part.state()
        np.testing.assert_allclose(sp["PSI"], sr["PSI"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(sp["Crs"], sr["Crs"], rtol=1e-12)
        np.testing.assert_allclose(sp["gr"], sr["gr"], rtol=1e-12)
    rr, rp = ref.finalize(), part.finalize()
    for k in ("FE", "CVBayes", "CVGP", "CVGT", "CVMAP", "varCVBayes"):
        np.testing.assert_allclose(getattr(rp, k), getattr(rr, k), rtol=1e-11, err_msg=k)
    res = part.em(itrmax=400)
    assert res.itrnum > 0 or not res.cnv
    part.close()
    ref.close()


def test_single_rank_partition_mod(gpu_required, single_rank_group):
    from graphbix_b200.engine import ModEngine
    from graphbix_b200.engine_dist import PartEngine
    q = 3
    n, links, psi0, _ = planted_case(2000, q, 9.0, 0.1, seed=77)
    gr = np.ones(q) / q
    ref = ModEngine(n, links, q)
    ref.init(gr, psi0)
    part = PartEngine(n, links, q, "mod")
    part.init(gr, None, psi0)
    for _ in range(5):
        ref.iterate(1)
        part.iterate(1)
        sr, sp = ref.state(), part.state()
        np.testing.assert_allclose(sp["PSI"], sr["PSI"], rtol=0, atol=1e-13)
        np.testing.assert_allclose([sp["omegain"], sp["omegaout"]], [sr["omegain"], sr["omegaout"]], rtol=1e-12)
    rr, rp = ref.finalize(), part.finalize()
    np.testing.assert_allclose(rp.FE, rr.FE, rtol=1e-11)
    part.close()
    ref.close()


def test_two_rank_partition_against_single_gpu(gpu_required):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run scripts/gpu_cycle.sh part on a multi-GPU box)")
    env = dict(os.environ)
    env.pop("RANK", None)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29578",
                          os.path.join(ROOT, "tests", "dist_check_part.py")], capture_output=True, text=True, env=env,
                         timeout=900)
    assert out.returncode == 0 and "PARTITIONED CHECK PASSED" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
