"""One strain matrix sharded over several GPUs (BASELINE config 4).

Input arrives channel-sharded: rank g holds channels [g*nx/G, (g+1)*nx/G) x all ns samples.  The f-k
filter couples all channels per time sample (column passes P1/P5) and all samples per kept
wavenumber row (time passes P2-P4), so the schedule is

    x (channel shards) --A2A--> time slabs --P1--> W slabs --A2A--> row shards --P2,P3,P4-->
                       <--A2A-- time slabs <--P5-- W slabs <--A2A--

Two of the four exchanges move the *pruned* spectrum only (1 356 of 5 001 rows for the fan mask), the
other two move the real matrix.  The kept rows are dealt to the ranks in contiguous runs of
`rows_per` (the last rank may get fewer): uneven all-to-all splits, no padding rows.

The filter is written as a sequence of local stages separated by exchanges (`stage_*`), driven either by
`torch.distributed.all_to_all_single` (NCCL over NVLink on GPUs, gloo in the CPU test) or, for tests on a single GPU,
by `run_local_group`, which steps G rank objects through the same stages in one process with an in-memory exchange --
so the real CUDA backend (d4w_fk_apply_pass_ex with time-slab / row-range geometry) is exercised without G GPUs.
"""
import numpy as np


def partition(nx, ns, nrows, world):
    """Static partition: channels per rank, samples per time slab, kept rows per rank."""
    if nx % world or ns % world:
        raise ValueError(f"nx={nx} and ns={ns} must be divisible by the number of ranks ({world})")
    rows_per = (nrows + world - 1) // world
    counts = [max(0, min(rows_per, nrows - r * rows_per)) for r in range(world)]
    return {"cpr": nx // world, "slab": ns // world, "rows_per": rows_per, "counts": counts}


class CudaBackend:
    """Compute steps on the local GPU through libd4w.so."""

    def __init__(self, mask, nx, ns, world, device=None, eps=0.0):
        import torch
        from . import _lib, fk
        self.torch, self._lib = torch, _lib
        self.device = torch.cuda.current_device() if device is None else device
        self.full = fk.get_plan(nx, ns, self.device)
        self.slab = fk.get_plan(nx, ns // world, self.device)
        self.dm = fk.device_mask_for(mask, self.full, eps)
        self.rows = self.dm.rows
        self.nx, self.ns, self.world = nx, ns, world
        with torch.cuda.device(self.device):
            # workspace of the column passes on a time slab: the kept rows W [rows][slab] followed by the level-A/B ring
            self.slab_ws_bytes = int(_lib.lib().d4w_fk_workspace_bytes(self.slab.ptr, self.dm.ptr))
        self._slab_ws = None

    def empty(self, shape, complex_=False):
        t = self.torch
        return t.empty(shape, dtype=t.complex64 if complex_ else t.float32, device=f"cuda:{self.device}")

    def slab_workspace(self):
        """(raw workspace, view of its first rows*slab complex values as [rows, slab])"""
        t = self.torch
        if self._slab_ws is None:
            self._slab_ws = t.empty(self.slab_ws_bytes, dtype=t.uint8, device=f"cuda:{self.device}")
        slab = self.ns // self.world
        w = self._slab_ws[: self.rows * slab * 8].view(t.complex64).view(self.rows, slab)
        return self._slab_ws, w

    def _pass(self, plan, x, y, ws, taper, p, s0, cnt, toff):
        L, lib = self._lib.lib(), self._lib
        null = lib.ffi.NULL
        with self.torch.cuda.device(self.device):
            lib.check(L.d4w_fk_apply_pass_ex(plan.ptr, self.dm.ptr, lib.ptr(x, "float*") if x is not None else null,
                                             lib.ptr(y, "float*") if y is not None else null, lib.ptr(ws), int(taper), p, s0, cnt,
                                             toff, lib.stream_ptr()), f"sharded fk pass {p}")

    def col_fwd(self, xs, taper, t_offset):                   # xs [nx, slab] -> W slab [rows, slab] complex
        ws, w = self.slab_workspace()
        self._pass(self.slab, xs, None, ws, taper, 1, 0, self.rows, t_offset)
        return w

    def row_filter(self, w_rows, slot_begin, count):          # w_rows [count, ns] complex, in place
        for p in (2, 3, 4):
            self._pass(self.full, None, None, w_rows, 0, p, slot_begin, count, 0)

    def col_inv_input(self):                                  # where the received W slab must be placed
        return self.slab_workspace()[1]

    def col_inv(self, ys):                                    # W slab (in the slab workspace) -> ys [nx, slab]
        ws, _ = self.slab_workspace()
        self._pass(self.slab, None, ys, ws, 0, 5, 0, self.rows, 0)


class ShardedFkFilter:
    """Rank-local part of the channel-sharded f-k filter.  `group=None` + torch.distributed initialised: collective mode;
    `rank`/`world` given explicitly: a member of an in-process group driven by run_local_group."""

    def __init__(self, nx, ns, backend, group=None, rank=None, world=None):
        self.be, self.group = backend, group
        if rank is None:
            import torch.distributed as dist
            self.dist = dist
            world, rank = dist.get_world_size(group), dist.get_rank(group)
        self.world, self.rank = world, rank
        self.nx, self.ns = nx, ns
        self.part = partition(nx, ns, backend.rows, self.world)

    # ---- local stages; each returns (send buffer, input split sizes, output split sizes) -------------------------------
    def stage0_pack_x(self, x_local):
        G, p = self.world, self.part
        assert tuple(x_local.shape) == (p["cpr"], self.ns)
        # channel shards -> time slabs: block j of the send buffer is x_local[:, slab_j]
        send = x_local.reshape(p["cpr"], G, p["slab"]).permute(1, 0, 2).contiguous()
        return send.reshape(G * p["cpr"], p["slab"]), None, None

    def stage1_col_fwd(self, recv, tapering):
        p = self.part
        xs = recv.reshape(self.nx, p["slab"])                      # blocks arrive in rank order = channel order
        w = self.be.col_fwd(xs, tapering, self.rank * p["slab"])    # [rows, slab], rows dealt to ranks in runs
        return w, list(p["counts"]), [p["counts"][self.rank]] * self.world

    def stage2_row_filter(self, recv):
        G, p = self.world, self.part
        cnt = p["counts"][self.rank]
        # received [G][cnt][slab] -> rows with a contiguous time axis [cnt][ns]
        w_rows = recv.reshape(G, cnt, p["slab"]).permute(1, 0, 2).contiguous().reshape(cnt, self.ns)
        if cnt:
            self.be.row_filter(w_rows, self.rank * p["rows_per"], cnt)
        send = w_rows.reshape(cnt, G, p["slab"]).permute(1, 0, 2).contiguous().reshape(G * cnt, p["slab"])
        return send, [cnt] * G, list(p["counts"])

    def stage3_recv_buffer(self):
        return self.be.col_inv_input()                             # [rows, slab] inside the column workspace

    def stage3_col_inv(self):
        p = self.part
        ys = self.be.empty((self.nx, p["slab"]))
        self.be.col_inv(ys)
        return ys, None, None                                      # [G*cpr, slab]: block j goes to rank j

    def stage4_unpack_y(self, recv):
        G, p = self.world, self.part
        return recv.reshape(G, p["cpr"], p["slab"]).permute(1, 0, 2).contiguous().reshape(p["cpr"], self.ns)

    # ---- collective driver ------------------------------------------------------------------------------------------
    def _a2a(self, send, in_split, out_split, out=None):
        if out is None:
            n_out = sum(out_split) if out_split is not None else send.shape[0]
            out = send.new_empty((n_out,) + tuple(send.shape[1:]))
        self.dist.all_to_all_single(out, send, output_split_sizes=out_split, input_split_sizes=in_split, group=self.group)
        return out

    def __call__(self, x_local, tapering=False, timers=None):
        """x_local: [nx/G, ns] float32 (this rank's channels). Returns the filtered [nx/G, ns].
        timers: optional dict; device milliseconds per stage are ADDED to it (CUDA events; one synchronize at the end)."""
        marks = []

        def mark(name):
            if timers is not None:
                import torch
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))
        mark("start")
        send = self.stage0_pack_x(x_local); mark("pack_x")
        recv = self._a2a(*send); mark("a2a_x_to_slabs")
        send = self.stage1_col_fwd(recv, tapering); mark("col_fwd")
        recv = self._a2a(*send); mark("a2a_w_to_rows")
        send, i_s, o_s = self.stage2_row_filter(recv); mark("row_filter_and_permutes")
        self._a2a(send, i_s, o_s, out=self.stage3_recv_buffer()); mark("a2a_w_to_slabs")
        send = self.stage3_col_inv(); mark("col_inv")
        recv = self._a2a(*send); mark("a2a_y_to_channels")
        y = self.stage4_unpack_y(recv); mark("unpack_y")
        if timers is not None:
            import torch
            torch.cuda.synchronize()
            for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
                timers[name] = timers.get(name, 0.0) + a.elapsed_time(b)
        return y


def _exchange(sends, in_splits, out_bufs=None):
    """In-process all-to-all: sends[r] is rank r's send buffer, split along dim 0 by in_splits[r] (None: equal parts)."""
    G = len(sends)
    parts = []
    for r in range(G):
        s = sends[r]
        sizes = in_splits[r] if in_splits[r] is not None else [s.shape[0] // G] * G
        offs = np.concatenate(([0], np.cumsum(sizes)))
        parts.append([s[offs[j]:offs[j + 1]] for j in range(G)])
    outs = []
    for r in range(G):
        blocks = [parts[j][r] for j in range(G)]
        if out_bufs is not None:
            o = 0
            for b in blocks:
                out_bufs[r][o:o + b.shape[0]].copy_(b)
                o += b.shape[0]
            outs.append(out_bufs[r])
        elif hasattr(blocks[0], "new_empty"):
            import torch
            outs.append(torch.cat(blocks, dim=0))
        else:
            outs.append(np.concatenate(blocks, axis=0))
    return outs


def run_local_group(filters, x_shards, tapering=False):
    """Drive G rank objects (one ShardedFkFilter each, any backend) through the sharded schedule in ONE process.
    Stages that share device state between "ranks" (the slab workspace of a backend object) run rank by rank."""
    G = len(filters)
    st = [f.stage0_pack_x(x) for f, x in zip(filters, x_shards)]
    recv = _exchange([s[0] for s in st], [s[1] for s in st])
    # each rank's W slab lives in its backend's workspace: copy it out before the next rank re-uses the same backend
    st = []
    for f, r in zip(filters, recv):
        w, i_s, o_s = f.stage1_col_fwd(r, tapering)
        st.append((w.clone() if hasattr(w, "clone") else w.copy(), i_s, o_s))
    recv = _exchange([s[0] for s in st], [s[1] for s in st])
    st = [f.stage2_row_filter(r) for f, r in zip(filters, recv)]
    slabs = _exchange([s[0] for s in st], [s[1] for s in st])          # per rank: [rows, slab]
    st = []
    for f, wslab in zip(filters, slabs):
        f.stage3_recv_buffer().copy_(wslab) if hasattr(wslab, "clone") else np.copyto(f.stage3_recv_buffer(), wslab)
        st.append(f.stage3_col_inv())
    recv = _exchange([s[0] for s in st], [s[1] for s in st])
    return [f.stage4_unpack_y(r) for f, r in zip(filters, recv)]


def fk_filter_filt_sharded(x_local, mask, nx, group=None, tapering=False, eps=0.0):
    """Channel-sharded dsp.fk_filter_filt: every rank passes its [nx/G, ns] float32 CUDA block."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    be = CudaBackend(mask, nx, x_local.shape[1], world, device=x_local.device.index, eps=eps)
    return ShardedFkFilter(nx, x_local.shape[1], be, group)(x_local, tapering=tapering)
