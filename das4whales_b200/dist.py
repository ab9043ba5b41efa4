"""One strain matrix sharded over several GPUs (BASELINE config 4).

Input arrives channel-sharded: rank g holds channels [g*nx/G, (g+1)*nx/G) x all ns samples.  The f-k
filter couples all channels per time sample (column passes P1/P5) and all samples per kept
wavenumber row (time passes P2-P4), so the schedule is

    x (channel shards) --A2A--> time slabs --P1--> W slabs --A2A--> row shards --P2,P3,P4-->
                       <--A2A-- time slabs <--P5-- W slabs <--A2A--

Two of the four exchanges move the *pruned* spectrum only (1 356 of 5 001 rows for the fan mask), the
other two move the real matrix.  The kept rows are dealt to the ranks in contiguous runs of
`rows_per` (the last rank may get fewer): uneven all-to-all splits, no padding rows.

Each rank's time slab is further cut into `nsub` sub-slabs: the exchanges are asynchronous collectives issued one sub-slab
ahead, so NVLink transfers run under the column transforms of the neighbouring sub-slab (SURVEY 7 hard part 6).

The filter is written as a sequence of local stages separated by exchanges (`stage_*`), driven either by
`torch.distributed.all_to_all_single` (NCCL over NVLink on GPUs, gloo in the CPU test) or, for tests on a single GPU,
by `run_local_group`, which steps G rank objects through the same stages in one process with an in-memory exchange --
so the real CUDA backend (d4w_fk_apply_pass_ex with time-slab / row-range geometry) is exercised without G GPUs.
"""
import numpy as np


def partition(nx, ns, nrows, world, nsub=1):
    """Static partition: channels per rank, samples per time slab / sub-slab, kept rows per rank."""
    if nx % world or ns % world:
        raise ValueError(f"nx={nx} and ns={ns} must be divisible by the number of ranks ({world})")
    if (ns // world) % nsub:
        raise ValueError(f"the time slab of {ns // world} samples is not divisible into {nsub} sub-slabs")
    rows_per = (nrows + world - 1) // world
    counts = [max(0, min(rows_per, nrows - r * rows_per)) for r in range(world)]
    return {"cpr": nx // world, "slab": ns // world, "sub": ns // world // nsub, "nsub": nsub, "rows_per": rows_per, "counts": counts}


def pick_nsub(ns, world, want=4):
    """Largest number of sub-slabs <= want whose length keeps the column kernels' alignment (multiple of 16 samples)."""
    slab = ns // world
    for n in range(want, 1, -1):
        if slab % n == 0 and (slab // n) % 16 == 0 and slab // n >= 4096:
            return n
    return 1


class CudaBackend:
    """Compute steps on the local GPU through libd4w.so.  The time slab of a rank is processed in `nsub` sub-slabs so that
    the exchange of one sub-slab overlaps with the column transform of another."""

    def __init__(self, mask, nx, ns, world, device=None, eps=0.0, nsub=1):
        import torch
        from . import _lib, fk
        self.torch, self._lib = torch, _lib
        self.device = torch.cuda.current_device() if device is None else device
        self.nsub = int(nsub)
        self.sub = ns // world // self.nsub
        self.full = fk.get_plan(nx, ns, self.device)
        self.slab = fk.get_plan(nx, self.sub, self.device)            # geometry of one column pass: [nx, sub]
        self.dm = fk.device_mask_for(mask, self.full, eps)
        self.rows = self.dm.rows
        self.nx, self.ns, self.world = nx, ns, world
        with torch.cuda.device(self.device):
            # workspace of a column pass: the kept rows W [rows][sub] followed by the level-A/B ring
            self.ws_bytes = int(_lib.lib().d4w_fk_workspace_bytes(self.slab.ptr, self.dm.ptr))
        self._ws = [None] * self.nsub

    def empty(self, shape, complex_=False):
        t = self.torch
        return t.empty(shape, dtype=t.complex64 if complex_ else t.float32, device=f"cuda:{self.device}")

    def workspace(self, s):
        """(raw workspace of sub-slab s, view of its first rows*sub complex values as [rows, sub])"""
        t = self.torch
        if self._ws[s] is None:
            self._ws[s] = t.empty(self.ws_bytes, dtype=t.uint8, device=f"cuda:{self.device}")
        w = self._ws[s][: self.rows * self.sub * 8].view(t.complex64).view(self.rows, self.sub)
        return self._ws[s], w

    def _pass(self, plan, x, y, ws, taper, p, s0, cnt, toff):
        L, lib = self._lib.lib(), self._lib
        null = lib.ffi.NULL
        with self.torch.cuda.device(self.device):
            lib.check(L.d4w_fk_apply_pass_ex(plan.ptr, self.dm.ptr, lib.ptr(x, "float*") if x is not None else null,
                                             lib.ptr(y, "float*") if y is not None else null, lib.ptr(ws), int(taper), p, s0, cnt,
                                             toff, lib.stream_ptr()), f"sharded fk pass {p}")

    def col_fwd(self, xs, taper, t_offset, s=0):               # xs [nx, sub] -> W [rows, sub] complex (workspace s)
        ws, w = self.workspace(s)
        self._pass(self.slab, xs, None, ws, taper, 1, 0, self.rows, t_offset)
        return w

    def row_filter(self, w_rows, slot_begin, count):           # w_rows [count, ns] complex, in place
        for p in (2, 3, 4):
            self._pass(self.full, None, None, w_rows, 0, p, slot_begin, count, 0)

    def col_inv_input(self, s=0):                              # where the received W sub-slab must be placed
        return self.workspace(s)[1]

    def col_inv(self, ys, s=0):                                # W (workspace s) -> ys [nx, sub]
        ws, _ = self.workspace(s)
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
        self.part = partition(nx, ns, backend.rows, self.world, getattr(backend, "nsub", 1))
        self._w_rows = None

    # ---- local stages (s = sub-slab index); "send" tuples are (buffer, input split sizes, output split sizes) ----------------
    def stage0_pack_x(self, x_local, s=0):
        G, p = self.world, self.part
        assert tuple(x_local.shape) == (p["cpr"], self.ns)
        # channel shards -> time sub-slabs: block j of the send buffer is x_local[:, slab_j][:, sub_s]
        send = x_local.reshape(p["cpr"], G, p["nsub"], p["sub"])[:, :, s, :].permute(1, 0, 2).contiguous()
        return send.reshape(G * p["cpr"], p["sub"]), None, None

    def stage1_col_fwd(self, recv, tapering, s=0):
        p = self.part
        xs = recv.reshape(self.nx, p["sub"])                       # blocks arrive in rank order = channel order
        w = self.be.col_fwd(xs, tapering, self.rank * p["slab"] + s * p["sub"], s)      # [rows, sub], rows dealt in runs
        return w, list(p["counts"]), [p["counts"][self.rank]] * self.world

    def rows_buffer(self):
        p = self.part
        cnt = p["counts"][self.rank]
        if self._w_rows is None:
            self._w_rows = self.be.empty((cnt, self.ns), complex_=True)
        return self._w_rows

    def stage2_place_rows(self, recv, s=0):
        """received [G][cnt][sub] of sub-slab s -> its time positions in the row buffer [cnt][G][nsub][sub] = [cnt][ns]"""
        G, p = self.world, self.part
        cnt = p["counts"][self.rank]
        dst = self.rows_buffer().reshape(cnt, G, p["nsub"], p["sub"])[:, :, s, :]
        dst.copy_(recv.reshape(G, cnt, p["sub"]).permute(1, 0, 2))

    def stage2_row_filter(self):
        p = self.part
        cnt = p["counts"][self.rank]
        if cnt:
            self.be.row_filter(self.rows_buffer(), self.rank * p["rows_per"], cnt)

    def stage2_pack_rows(self, s=0):
        G, p = self.world, self.part
        cnt = p["counts"][self.rank]
        send = self.rows_buffer().reshape(cnt, G, p["nsub"], p["sub"])[:, :, s, :].permute(1, 0, 2).contiguous()
        return send.reshape(G * cnt, p["sub"]), [cnt] * G, list(p["counts"])

    def stage3_recv_buffer(self, s=0):
        return self.be.col_inv_input(s)                            # [rows, sub] inside workspace s

    def stage3_col_inv(self, s=0):
        p = self.part
        ys = self.be.empty((self.nx, p["sub"]))
        self.be.col_inv(ys, s)
        return ys, None, None                                      # [G*cpr, sub]: block j goes to rank j

    def stage4_unpack_y(self, recv, y_local, s=0):
        G, p = self.world, self.part
        dst = y_local.reshape(p["cpr"], G, p["nsub"], p["sub"])[:, :, s, :]
        dst.copy_(recv.reshape(G, p["cpr"], p["sub"]).permute(1, 0, 2))

    # ---- collective driver ------------------------------------------------------------------------------------------
    def _a2a(self, send, in_split, out_split, out=None, async_op=False):
        if out is None:
            n_out = sum(out_split) if out_split is not None else send.shape[0]
            out = send.new_empty((n_out,) + tuple(send.shape[1:]))
        work = self.dist.all_to_all_single(out, send, output_split_sizes=out_split, input_split_sizes=in_split, group=self.group,
                                           async_op=async_op)
        return (out, work, send) if async_op else out

    def __call__(self, x_local, tapering=False, timers=None):
        """x_local: [nx/G, ns] float32 (this rank's channels). Returns the filtered [nx/G, ns].

        With nsub > 1 the exchanges are asynchronous collectives issued one sub-slab ahead, so the transfer of sub-slab s+1
        runs under the column transform of sub-slab s (forward half) and the transfer of the filtered sub-slab s under the
        inverse transform of s+1 (inverse half).  timers (nsub = 1 only): device milliseconds per stage are ADDED to the dict."""
        S = self.part["nsub"]
        y_local = x_local.new_empty(x_local.shape)
        if S == 1:
            return self._call_serial(x_local, y_local, tapering, timers)
        # ---- forward half: x sub-slabs -> column transform -> kept rows to their owners
        pend_x = {}

        def issue_x(s):
            pend_x[s] = self._a2a(*self.stage0_pack_x(x_local, s), async_op=True)
        issue_x(0)
        if S > 1:
            issue_x(1)
        pend_w = []
        for s in range(S):
            recv, work, _keep = pend_x.pop(s)
            work.wait()
            pend_w.append(self._a2a(*self.stage1_col_fwd(recv, tapering, s), async_op=True))
            if s + 2 < S:
                issue_x(s + 2)
        for s, (recv, work, _keep) in enumerate(pend_w):
            work.wait()
            self.stage2_place_rows(recv, s)
        del pend_w
        self.stage2_row_filter()
        # ---- inverse half: rows back to time sub-slabs -> inverse column transform -> channels
        pend_r = {}

        def issue_r(s):
            send, i_s, o_s = self.stage2_pack_rows(s)
            pend_r[s] = self._a2a(send, i_s, o_s, out=self.stage3_recv_buffer(s), async_op=True)
        issue_r(0)
        if S > 1:
            issue_r(1)
        pend_y = []
        for s in range(S):
            _out, work, _keep = pend_r.pop(s)
            work.wait()
            pend_y.append(self._a2a(*self.stage3_col_inv(s), async_op=True))
            if s + 2 < S:
                issue_r(s + 2)
        for s, (recv, work, _keep) in enumerate(pend_y):
            work.wait()
            self.stage4_unpack_y(recv, y_local, s)
        return y_local

    def _call_serial(self, x_local, y_local, tapering, timers):
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
        self.stage2_place_rows(recv); self.stage2_row_filter(); send, i_s, o_s = self.stage2_pack_rows(); mark("row_filter_and_permutes")
        self._a2a(send, i_s, o_s, out=self.stage3_recv_buffer()); mark("a2a_w_to_slabs")
        send = self.stage3_col_inv(); mark("col_inv")
        recv = self._a2a(*send); mark("a2a_y_to_channels")
        self.stage4_unpack_y(recv, y_local); mark("unpack_y")
        if timers is not None:
            import torch
            torch.cuda.synchronize()
            for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
                timers[name] = timers.get(name, 0.0) + a.elapsed_time(b)
        return y_local


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
    """Drive G rank objects (one ShardedFkFilter each, any backend) through the sharded schedule in ONE process, sub-slab by
    sub-slab.  The rank objects may share one backend (one GPU): whatever lives in a backend workspace is copied out before
    the next "rank" re-uses it."""
    S = filters[0].part["nsub"]
    ys = [x.new_empty(x.shape) for x in x_shards]
    for s in range(S):
        st = [f.stage0_pack_x(x, s) for f, x in zip(filters, x_shards)]
        recv = _exchange([t[0] for t in st], [t[1] for t in st])
        st = []
        for f, r in zip(filters, recv):
            w, i_s, o_s = f.stage1_col_fwd(r, tapering, s)
            st.append((w.clone(), i_s, o_s))
        recv = _exchange([t[0] for t in st], [t[1] for t in st])
        for f, r in zip(filters, recv):
            f.stage2_place_rows(r, s)
    for f in filters:
        f.stage2_row_filter()
    for s in range(S):
        st = [f.stage2_pack_rows(s) for f in filters]
        slabs = _exchange([t[0] for t in st], [t[1] for t in st])          # per rank: [rows, sub]
        st = []
        for f, wsub in zip(filters, slabs):
            f.stage3_recv_buffer(s).copy_(wsub)
            st.append(f.stage3_col_inv(s))
        recv = _exchange([t[0] for t in st], [t[1] for t in st])
        for f, r, y in zip(filters, recv, ys):
            f.stage4_unpack_y(r, y, s)
    return ys


def fk_filter_filt_sharded(x_local, mask, nx, group=None, tapering=False, eps=0.0, nsub=None):
    """Channel-sharded dsp.fk_filter_filt: every rank passes its [nx/G, ns] float32 CUDA block."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    ns = x_local.shape[1]
    be = CudaBackend(mask, nx, ns, world, device=x_local.device.index, eps=eps, nsub=pick_nsub(ns, world) if nsub is None else nsub)
    return ShardedFkFilter(nx, ns, be, group)(x_local, tapering=tapering)
