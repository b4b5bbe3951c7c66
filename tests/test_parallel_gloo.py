"""T5 host-side logic of the data-parallel step on CPU (world size 2, gloo): frame sharding, the flat
gradient buffer layout and the sum all-reduce.  The compute on each rank is the oracle (tests may use
it); the product's CUDA step (`formation_step`) uses the same `shard_frames` / `GradLayout` / comm."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from casualhdrsplat_b200.parallel import GradLayout, TorchComm, shard_frames
from casualhdrsplat_b200.scene import make_scene
from tests.util import LEAF_NAMES


def test_shard_frames_partitions():
    for B in [1, 7, 8, 9]:
        for world in [1, 2, 3, 4, 8]:
            got = [i for r in range(world) for i in shard_frames(B, r, world)]
            assert got == list(range(B))
            sizes = [len(shard_frames(B, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_frames(4, 4, 4)


def test_grad_layout_views_are_disjoint_and_cover():
    lay = GradLayout(n_gauss=5, n_knots=6, n_crf=3 * 49, n_frames=4)
    buf = torch.zeros(lay.total)
    v = lay.views(buf)
    k = 1.0
    for name in ["means", "quats", "scales", "opacities", "colors", "knots", "crf_params", "exposure_times", "frame_times"]:
        v[name] += k
        k += 1
    assert (buf != 0).all() and buf.numel() == 14 * 5 + 42 + 147 + 8
    assert v["means"].shape == (5, 3) and v["quats"].shape == (5, 4) and v["knots"].shape == (6, 7)


def _oracle_grads(sc, frame_ids):
    import oracle

    leaves = {k: getattr(sc, k).double().requires_grad_(True) for k in LEAF_NAMES}
    idx = torch.tensor(list(frame_ids))
    sp = dict(knots=leaves["knots"], knot_t0=sc.knot_t0, knot_dt=sc.knot_dt, frame_times=leaves["frame_times"][idx], kind=sc.spline_kind)
    ldr, _, _ = oracle.rasterize(leaves["means"], leaves["quats"], leaves["scales"], leaves["opacities"], leaves["colors"], None,
                                 sc.Ks[idx], sc.width, sc.height, leaves["exposure_times"][idx], sc.n_virtual, sc.crf_kind,
                                 leaves["crf_params"], spline=sp)
    gs = torch.autograd.grad((ldr * sc.v_ldr[idx].double()).sum(), list(leaves.values()))
    return dict(zip(leaves.keys(), gs))


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        sc = make_scene(300, 48, 32, n_frames=3, n_virtual=2, crf_hidden=8, scale_mult=14.0)
        ids = shard_frames(sc.n_frames, rank, world)
        g = _oracle_grads(sc, ids)
        lay = GradLayout(sc.means.shape[0], sc.knots.shape[0], sc.crf_params.numel(), sc.n_frames)
        buf = torch.zeros(lay.total, dtype=torch.float64)
        v = lay.views(buf)
        for k in ["means", "quats", "scales", "opacities", "colors", "knots"]:
            v[k].copy_(g[k])
        v["crf_params"].copy_(g["crf_params"].reshape(-1))
        v["exposure_times"].copy_(g["exposure_times"])  # autograd already scatters into the frames this rank owns
        v["frame_times"].copy_(g["frame_times"])
        comm = TorchComm()
        assert comm.world == world and comm.rank == rank
        comm.allreduce_(buf)
        if rank == 0:
            torch.save(buf, out_path)
    finally:
        dist.destroy_process_group()


def test_sharded_gradients_sum_to_full_batch(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "reduced.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    buf = torch.load(out)
    sc = make_scene(300, 48, 32, n_frames=3, n_virtual=2, crf_hidden=8, scale_mult=14.0)
    full = _oracle_grads(sc, range(sc.n_frames))
    lay = GradLayout(sc.means.shape[0], sc.knots.shape[0], sc.crf_params.numel(), sc.n_frames)
    v = lay.views(buf)
    for k in ["means", "quats", "scales", "opacities", "colors", "knots", "exposure_times", "frame_times"]:
        assert torch.allclose(v[k], full[k], rtol=1e-9, atol=1e-12 * float(full[k].abs().max())), k
    assert torch.allclose(v["crf_params"], full["crf_params"].reshape(-1), rtol=1e-9, atol=1e-12)


def _params_worker(rank, world, port, out_dir):
    from casualhdrsplat_b200.parallel import ShardedHostParams

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        host = {"means": torch.randn(37, 3, generator=g), "opacities": torch.rand(37, generator=g), "crf": torch.randn(3, 7, generator=g),
                "scalar": torch.randn((), generator=g)}
        sp = ShardedHostParams(host, torch.device("cpu"), rank, world, comm=None, n_slots=2)
        # a rank only has to hold ITS slice on the host: wipe the rest to prove the exchange supplies it
        keep = sp.host_flat[rank * sp.per:(rank + 1) * sp.per].clone()
        sp.host_flat.zero_()
        sp.host_flat[rank * sp.per:(rank + 1) * sp.per] = keep
        for slot in range(2):
            sp.upload_(slot)
            v = sp.views(slot)
            for k in host:
                assert torch.equal(v[k], host[k]), (rank, slot, k)
        n = 14 * 37 + 11
        slices = [None] * world
        dist.all_gather_object(slices, sp.grad_slice(n))
        assert sp.h2d_bytes == sp.per * 4 and sp.per * world >= sum(v.numel() for v in host.values())
        if rank == 0:
            torch.save(slices, os.path.join(out_dir, "slices.pt"))
    finally:
        dist.destroy_process_group()


def test_sharded_host_params_replicate_and_gradient_slices_cover(tmp_path):
    """Each rank uploads 1/G of the flat parameter buffer and the exchange rebuilds every tensor on every rank; the gradient
    read-back slices of the ranks tile the flat gradient buffer exactly once."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_params_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    slices = torch.load(os.path.join(str(tmp_path), "slices.pt"))
    n = 14 * 37 + 11
    assert slices[0][0] == 0 and slices[-1][1] == n
    for a, b in zip(slices[:-1], slices[1:]):
        assert a[1] == b[0] and a[0] % 4 == 0
