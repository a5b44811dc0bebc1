"""CPU tier: the N>1 path (contiguous sharding + the single all-gather) on 2 gloo ranks, each rank solving its shard on
the test-only CTA emulator; the gathered result must equal the single-process solve of the whole batch bit for bit."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
import algames_b200 as ab
from algames_b200 import distributed as D
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
model, N, dt, obj, con, opts, x0, xf = ab.workloads.config_b(batch=5, N=10)
rng = np.random.default_rng(5)
Z0 = 1e-8 * rng.random((5, N, model.n + model.m)); L0 = 1e-8 * rng.random((5, model.p, N - 1, model.n))
mk = lambda B: ab.GameBatch(model, N, dt, obj, con, B, lib_path=%(emu)r)
res = D.solve_sharded(mk, x0, xf, Z0, L0, opts, rank, world)
if rank == 0:
    np.savez(%(out)r, **res)
dist.barrier()
dist.destroy_process_group()
'''


def test_sharded_solve_two_gloo_ranks(tmp_path):
    import test_emulated_kernels as T
    emu = T.EMU
    if not os.path.exists(emu):
        subprocess.run([os.path.join(HERE, "emu", "build.sh")], check=True)
    out = str(tmp_path / "gathered.npz")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "emu": emu, "out": out})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29631")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                    "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)], check=True, env=env, timeout=600)
    got = np.load(out)
    import algames_b200 as ab
    from algames_b200 import distributed as D
    assert D.shard_bounds(5, 2, 0) == (0, 3) and D.shard_bounds(5, 2, 1) == (3, 5)
    model, N, dt, obj, con, opts, x0, xf = ab.workloads.config_b(batch=5, N=10)          # 5 instances on 2 ranks: uneven shards (3 + 2)
    rng = np.random.default_rng(5)
    Z0 = 1e-8 * rng.random((5, N, model.n + model.m)); L0 = 1e-8 * rng.random((5, model.p, N - 1, model.n))
    ref = D.solve_sharded(lambda B: ab.GameBatch(model, N, dt, obj, con, B, lib_path=emu), x0, xf, Z0, L0, opts, 0, 1)
    for k in ("Z", "L", "stats", "status"):
        assert np.array_equal(got[k], ref[k]), k
    assert (ref["status"] == 0).all() and ref["Z"].shape[0] == 5
    import pytest
    with pytest.raises(ValueError):
        D.solve_sharded(None, x0[:1], None, Z0[:1], L0[:1], opts, 0, 2)
