"""world_size-2 gloo test of the multi-GPU logic (no GPU needed): each rank steps its shard of the
environments with the host emulation of the kernels; the gathered result must be bit-identical to one
process stepping the whole batch.  TEST INFRASTRUCTURE: hostemu is the same kernel source compiled for
the host (tests/hostemu), never used by the product."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from mujoco_b200.shard import shard_range  # noqa: E402


def test_shard_range_partitions():
    for n in [1, 7, 8, 4096, 4097]:
        for w in [1, 2, 3, 8]:
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _states_and_ctrl(nenv, nstep):
    from mjb_util import HUMANOID, perturbed_states
    from oracle_util import Oracle
    o = Oracle(HUMANOID)
    o.set_opt("solver", 0)
    st = perturbed_states(o, nenv, seed=11, height=[0.9, 1.3, 0.4], qvel_std=0.5, qpos_std=0.1)
    ctrl = np.random.default_rng(5).uniform(-1, 1, (nenv, nstep, o.size("nu")))
    return st, ctrl


def _run(rank, world, port, nenv, nstep, out_path):
    import torch
    import torch.distributed as dist
    import mujoco_b200 as mb
    from mjb_util import HUMANOID, hostemu_lib
    from mujoco_b200.shard import gather_env_rows, max_over_ranks
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        st, ctrl = _states_and_ctrl(nenv, nstep)
        lo, hi = shard_range(nenv, rank, world)
        m = mb.Model(HUMANOID, library=hostemu_lib())
        m.set_option("solver", mb.SOLVER_PGS)
        b = mb.Batch(m, hi - lo, nconmax=48, njmax=128)
        traj = b.rollout(st[lo:hi], ctrl[lo:hi])            # [n_local, nstep, nstate]
        final = gather_env_rows(traj[:, -1, :], nenv)
        tmax = max_over_ranks(1.0 + rank)
        assert tmax == float(world)
        if rank == 0:
            np.save(out_path, final.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_shards_match_single_process(tmp_path):
    import torch.multiprocessing as mp
    import mujoco_b200 as mb
    from mjb_util import HOSTEMU, HUMANOID, hostemu_lib
    if not os.path.exists(HOSTEMU):
        pytest.skip("host emulation library not built")
    nenv, nstep, world = 5, 12, 2       # odd count: shards of 3 and 2 envs
    out = str(tmp_path / "final.npy")
    mp.spawn(_run, args=(world, _free_port(), nenv, nstep, out), nprocs=world, join=True)
    got = np.load(out)
    st, ctrl = _states_and_ctrl(nenv, nstep)
    m = mb.Model(HUMANOID, library=hostemu_lib())
    m.set_option("solver", mb.SOLVER_PGS)
    b = mb.Batch(m, nenv, nconmax=48, njmax=128)
    want = b.rollout(st, ctrl)[:, -1, :]
    assert got.shape == want.shape
    assert np.array_equal(got, want)
