"""CPU, world size 2, gloo: the partition / all-to-all choreography of the channel-sharded f-k
filter (das4whales_b200/dist.py) with a NumPy stand-in for the three local compute steps.  The
result on every rank must equal the single-process float64 oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DX, FS = 2.0419046878814697, 200.0


class NumpyBackend:
    """Same contract as dist.CudaBackend, arithmetic in float64 NumPy (test stand-in only)."""

    def __init__(self, mask_shifted, nx, ns, world, nsub=1):
        from oracle import dsp_oracle as O
        self.nsub = nsub
        self._wsub = {}
        self.msym = O.fold_mask(mask_shifted)                    # un-shifted layout
        self.act = [k for k in range(nx // 2 + 1) if np.any(self.msym[k] != 0)]
        self.rows = len(self.act)
        self.nx, self.ns, self.world = nx, ns, world
        self.taper = O.tukey_window(ns)

    def empty(self, shape, complex_=False):
        return torch.zeros(shape, dtype=torch.complex128 if complex_ else torch.float64)

    def col_fwd(self, xs, taper, t_offset, s=0):
        x = xs.numpy()
        if taper:
            x = x * self.taper[t_offset:t_offset + x.shape[1]][None, :]
        return torch.from_numpy(np.fft.fft(x, axis=0)[self.act])

    def row_filter(self, w_rows, slot_begin, count):
        if count:
            ks = self.act[slot_begin:slot_begin + count]
            f = np.fft.fft(w_rows[:count].numpy(), axis=1) * self.msym[ks]
            w_rows[:count] = torch.from_numpy(np.fft.ifft(f, axis=1))

    def col_inv_input(self, s=0):
        if s not in self._wsub:
            self._wsub[s] = torch.zeros((self.rows, self.ns // self.world // self.nsub), dtype=torch.complex128)
        return self._wsub[s]

    def col_inv(self, ys, s=0):
        w = self._wsub[s].numpy()
        spec = np.zeros((self.nx, w.shape[1]), dtype=np.complex128)
        for s, k in enumerate(self.act):
            spec[k] = w[s]
            if k != 0 and 2 * k != self.nx:
                spec[self.nx - k] = np.conj(w[s])
        ys[:] = torch.from_numpy(np.fft.ifft(spec, axis=0).real)


def _worker(rank, world, port, nx, ns, taper, q, nsub=1):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from das4whales_b200 import dist as d4wdist
    from oracle import dsp_oracle as O
    rng = np.random.default_rng(0)
    x = rng.standard_normal((nx, ns))
    mask = O.fk_filter_design((nx, ns), [0, nx, 1], DX, FS)
    be = NumpyBackend(mask, nx, ns, world, nsub)
    flt = d4wdist.ShardedFkFilter(nx, ns, be)
    cpr = nx // world
    y_local = flt(torch.from_numpy(x[rank * cpr:(rank + 1) * cpr].copy()), tapering=taper).numpy()
    ref = O.fk_filter_filt(x.copy(), mask, tapering=taper)[rank * cpr:(rank + 1) * cpr]
    err = float(np.max(np.abs(y_local - ref)) / np.max(np.abs(ref)))
    q.put((rank, err, flt.part["rows_per"], be.rows))
    dist.destroy_process_group()


@pytest.mark.parametrize("nx,ns,taper,nsub", [(24, 160, False, 1), (30, 96, True, 1), (24, 160, True, 4), (30, 96, False, 3)])
def test_sharded_fk_world2_gloo(nx, ns, taper, nsub):
    """nsub > 1 runs the pipelined driver: asynchronous all-to-alls issued one sub-slab ahead of the compute steps"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + nx + 7 * nsub) % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nx, ns, taper, q, nsub)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, rows_per, rows in res:
        assert err <= 1e-12, (rank, err)
        assert rows_per * 2 >= rows > 0


@pytest.mark.parametrize("nx,ns,world,nsub", [(24, 160, 2, 1), (36, 96, 3, 2), (32, 64, 4, 1), (24, 160, 2, 4)])
def test_local_group_driver_matches_oracle(nx, ns, world, nsub):
    """dist.run_local_group (the in-process stand-in for the collectives, used by the single-GPU test of the CUDA
    backend) against the float64 oracle, uneven row splits included."""
    sys.path.insert(0, ROOT)
    from das4whales_b200 import dist as d4wdist
    from oracle import dsp_oracle as O
    rng = np.random.default_rng(1)
    x = rng.standard_normal((nx, ns))
    mask = O.fk_filter_design((nx, ns), [0, nx, 1], DX, FS)
    cpr = nx // world
    filters = [d4wdist.ShardedFkFilter(nx, ns, NumpyBackend(mask, nx, ns, world, nsub), rank=r, world=world) for r in range(world)]
    ys = d4wdist.run_local_group(filters, [torch.from_numpy(x[r * cpr:(r + 1) * cpr].copy()) for r in range(world)], tapering=True)
    ref = O.fk_filter_filt(x.copy(), mask, tapering=True)
    got = np.concatenate([y.numpy() for y in ys], axis=0)
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) <= 1e-12


def test_partition_rules():
    sys.path.insert(0, ROOT)
    from das4whales_b200.dist import partition
    p = partition(20000, 240000, 2711, 4)
    assert p == {"cpr": 5000, "slab": 60000, "sub": 60000, "nsub": 1, "rows_per": 678, "counts": [678, 678, 678, 677]}
    assert partition(20000, 240000, 2711, 4, 4)["sub"] == 15000
    from das4whales_b200.dist import pick_nsub
    assert pick_nsub(240000, 4) == 3 and pick_nsub(240000, 2) == 4 and pick_nsub(4800, 2) == 1
    assert partition(8, 16, 3, 4)["counts"] == [1, 1, 1, 0]
    with pytest.raises(ValueError):
        partition(10001, 120000, 100, 2)
