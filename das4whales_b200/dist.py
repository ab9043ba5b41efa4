"""One strain matrix sharded over several GPUs (BASELINE config 4).

Input arrives channel-sharded: rank g holds channels [g*nx/G, (g+1)*nx/G) x all ns samples.  The f-k
filter couples all channels per time sample (column passes P1/P5) and all samples per kept
wavenumber row (time passes P2-P4), so the schedule is

    x (channel shards) --A2A--> time slabs --P1--> W slabs --A2A--> row shards --P2,P3,P4-->
                       <--A2A-- time slabs <--P5-- W slabs <--A2A--

Two of the four exchanges move the *pruned* spectrum only (1 356 of 5 001 rows for the fan mask), the
other two move the real matrix.  Exchanges are torch.distributed all_to_all_single (NCCL over
NVLink on GPUs, gloo in the CPU test); compute steps come from a backend object so the partition /
exchange logic is testable without a GPU (tests/test_dist_gloo.py).
"""
import numpy as np


def partition(nx, ns, nrows, world):
    """Static partition: channels per rank, samples per time slab, kept rows per rank (padded)."""
    if nx % world or ns % world:
        raise ValueError(f"nx={nx} and ns={ns} must be divisible by the number of ranks ({world})")
    rows_per = (nrows + world - 1) // world
    return {"cpr": nx // world, "slab": ns // world, "rows_per": rows_per, "rows_pad": rows_per * world}


class CudaBackend:
    """Compute steps on the local GPU through libd4w.so."""

    def __init__(self, mask, nx, ns, world, device=None):
        import torch
        from . import _lib, fk
        self.torch, self._lib = torch, _lib
        self.device = torch.cuda.current_device() if device is None else device
        self.full = fk.get_plan(nx, ns, self.device)
        self.slab = fk.get_plan(nx, ns // world, self.device)
        self.dm = fk.device_mask_for(mask, self.full)
        self.rows = self.dm.rows
        self.nx, self.ns, self.world = nx, ns, world

    def empty(self, shape, complex_=False):
        t = self.torch
        return t.empty(shape, dtype=t.complex64 if complex_ else t.float32, device=f"cuda:{self.device}")

    def _pass(self, plan, x, y, ws, taper, p, s0, cnt, toff):
        L, lib = self._lib.lib(), self._lib
        null = lib.ffi.NULL
        lib.check(L.d4w_fk_apply_pass_ex(plan.ptr, self.dm.ptr, lib.ptr(x, "float*") if x is not None else null,
                                         lib.ptr(y, "float*") if y is not None else null, lib.ptr(ws), int(taper), p, s0, cnt,
                                         toff, lib.stream_ptr()), f"sharded fk pass {p}")

    def col_fwd(self, xs, w_slab, taper, t_offset):           # xs [nx, slab] -> w_slab [rows_pad, slab] complex
        self._pass(self.slab, xs, None, w_slab, taper, 1, 0, self.rows, t_offset)

    def row_filter(self, w_rows, slot_begin, count):          # w_rows [rows_per, ns] complex, in place
        for p in (2, 3, 4):
            self._pass(self.full, None, None, w_rows, 0, p, slot_begin, count, 0)

    def col_inv(self, w_slab, ys):                            # w_slab -> ys [nx, slab]
        self._pass(self.slab, None, ys, w_slab, 0, 5, 0, self.rows, 0)


class ShardedFkFilter:
    def __init__(self, nx, ns, backend, group=None):
        import torch.distributed as dist
        self.dist, self.group, self.be = dist, group, backend
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.nx, self.ns = nx, ns
        self.part = partition(nx, ns, backend.rows, self.world)

    def _a2a(self, src):
        dst = src.new_empty(src.shape)
        self.dist.all_to_all_single(dst, src.contiguous(), group=self.group)
        return dst

    def __call__(self, x_local, tapering=False):
        """x_local: [nx/G, ns] float32 (this rank's channels). Returns the filtered [nx/G, ns]."""
        G, p, be = self.world, self.part, self.be
        cpr, slab, rp = p["cpr"], p["slab"], p["rows_per"]
        assert tuple(x_local.shape) == (cpr, self.ns)
        # channel shards -> time slabs: block j of the send buffer is x_local[:, slab_j]
        xs = self._a2a(x_local.reshape(cpr, G, slab).permute(1, 0, 2)).reshape(self.nx, slab)
        w_slab = be.empty((p["rows_pad"], slab), complex_=True)
        if p["rows_pad"] > be.rows:
            w_slab[be.rows:] = 0
        be.col_fwd(xs, w_slab, tapering, self.rank * slab)
        # time slabs -> row shards: block j of the send buffer is w_slab[rows_j, :]
        w_rows = self._a2a(w_slab.reshape(G, rp, slab)).permute(1, 0, 2).reshape(rp, self.ns).contiguous()
        s0 = self.rank * rp
        cnt = max(0, min(rp, be.rows - s0))
        be.row_filter(w_rows, s0, cnt)
        w_slab = self._a2a(w_rows.reshape(rp, G, slab).permute(1, 0, 2)).reshape(p["rows_pad"], slab)
        ys = be.empty((self.nx, slab))
        be.col_inv(w_slab, ys)
        return self._a2a(ys.reshape(G, cpr, slab)).permute(1, 0, 2).reshape(cpr, self.ns).contiguous()


def fk_filter_filt_sharded(x_local, mask, nx, group=None, tapering=False):
    """Channel-sharded dsp.fk_filter_filt: every rank passes its [nx/G, ns] float32 CUDA block."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    be = CudaBackend(mask, nx, x_local.shape[1], world, device=x_local.device.index)
    return ShardedFkFilter(nx, x_local.shape[1], be, group)(x_local, tapering=tapering)
