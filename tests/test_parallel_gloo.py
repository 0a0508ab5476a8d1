"""CPU, world_size 2 over gloo: the rank-parallel q sweep deals tasks round-robin and gathers them in order."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_round_robin_share():
    from graphbix_b200.parallel import my_share
    tasks = list(range(2, 11))
    a, b = my_share(tasks, 0, 2), my_share(tasks, 1, 2)
    assert [t for _, t in a] == [2, 4, 6, 8, 10] and [t for _, t in b] == [3, 5, 7, 9]
    assert sorted(i for i, _ in a + b) == list(range(len(tasks)))


def test_sweep_two_ranks_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, json
        sys.path.insert(0, {ROOT!r})
        from graphbix_b200.parallel import sweep
        res = sweep(range(2, 11), lambda q, dev: dict(q=q, rank=int(os.environ["RANK"]), sq=q * q), backend="gloo")
        if int(os.environ["RANK"]) == 0:
            print("RESULT " + json.dumps(res))
        else:
            assert res is None
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    import json
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][0]
    res = json.loads(line[7:])
    assert [r["q"] for r in res] == list(range(2, 11))
    assert [r["rank"] for r in res] == [i % 2 for i in range(9)]
    assert all(r["sq"] == r["q"] ** 2 for r in res)
