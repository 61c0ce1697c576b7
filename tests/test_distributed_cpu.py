"""Multi-rank host logic on CPU: world_size-2 (and 3) gloo process groups exercise the pixel sharding, the
convergence-norm all-reduce and the output gather that the N-GPU path uses with nccl.  The per-slab compute is
replaced by the float64 oracle here (no GPU in this test); the sharded result must equal the single-process one."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from csromer_b200 import distributed as cd


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 262144, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [cd.shard_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        cd.shard_bounds(10, 2, 2)


def test_single_process_paths():
    x = torch.arange(12.0).reshape(3, 4)
    local, (lo, hi) = cd.shard_pixels(x)
    assert (lo, hi) == (0, 4) and torch.equal(local, x)
    assert cd.global_convergence_norm(torch.tensor(6.0), 3) == 2.0
    assert cd.max_over_ranks(1.5) == 1.5
    assert torch.equal(cd.gather_pixels(x, 4), x)


def _worker(rank, world, port, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle import restatement as ob
        g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mixed_noise1_jvla200_k60.npz"))
        rng = np.random.RandomState(0)  # same cube on every rank
        Y, X = 3, 5
        scale = rng.uniform(0.5, 2.0, size=(Y, X))
        cube = g["data"][:, None, None] * scale[None]
        op = ob.ExactDFT(g["lambda2"], g["phi"], g["w"], g["s"], float(g["k"]))
        lam0, noise, K = float(g["lam0"]), float(g["noise"]), 12
        scale_flat = scale.reshape(-1)

        def make_problem(slab):
            return slab

        state = {"seen": 0}

        def run_slab(slab):
            lo = state["lo"] + state["seen"]
            sc = scale_flat[lo:lo + slab.shape[1]]
            state["seen"] += slab.shape[1]
            x0 = ob.complex_to_real(op.backward(slab))
            r = ob.fista(op, slab, x0, lam0 * sc, noise * sc, K)
            return torch.as_tensor(r["x"]), float(np.abs(r["x"]).sum())

        state["lo"] = cd.shard_bounds(Y * X, world, rank)[0]
        x_full, norm, (lo, hi) = cd.reconstruct_cube(make_problem, run_slab, cube, pixel_chunk=4)
        assert x_full.shape == (2 * len(g["phi"]), Y * X)
        # single-process reference
        flat = cube.reshape(cube.shape[0], -1)
        ref = ob.fista(op, flat, ob.complex_to_real(op.backward(flat)), lam0 * scale_flat, noise * scale_flat, K)
        assert np.abs(x_full.numpy() - ref["x"]).max() <= 1e-12 * np.abs(ref["x"]).max()
        want = np.abs(ref["x"]).sum() / (Y * X)
        assert abs(norm - want) <= 1e-9 * want, (norm, want)
        assert cd.max_over_ranks(float(rank)) == world - 1
        # gather to one rank only
        local = x_full[:, lo:hi].contiguous()
        got = cd.gather_pixels(local, Y * X, dst=0)
        if rank == 0:
            assert torch.equal(got, x_full)
        else:
            assert got is None
        open(os.path.join(tmpdir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_cube_equals_single_process(world, tmp_path):
    port = 29500 + (os.getpid() % 400) + world
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))
