"""One process per GPU: the colour set of a Jacobian is partitioned over the ranks.

Colours are independent units given x (each colour = perturb -> f! -> diff -> scatter into a disjoint set of J slots),
so there is NO collective on the data path until the end, where every rank needs the entries the others computed.
Two implementations of that final exchange:

  "p2p"  (default on NVLink/NVSwitch)  the diff+scatter kernel itself stores every value it owns into rank 0's nzval
         buffer (gather="root": rank 0's buffer is mapped by every rank over CUDA IPC and passed as THE J of plans created
         with shared_j) — compute and gather are one kernel; the ranks are ordered by the C ABI's device-side barrier
         (fdb_sync, DeviceBarrier below), no NCCL call per Jacobian.  gather="all" adds one NCCL broadcast of nzval;
         gather="all_p2p" stores to every peer from the kernel (fine up to 4 GPUs, pathological at 8 — see ShardedJacobian).
  "nccl" (fallback; also what the CPU/gloo tests exercise)  each rank packs the entries it owns into a compact
         buffer, all_gather, then un-permutes into nzval.

torch.distributed is plumbing only (rendezvous, barrier, the fallback all_gather); the hot path is libfdjac_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L
from . import api


# ------------------------------------------------------------------------------------------------ host-side partition logic
def partition_colors(n_colors: int, world: int, counts=None, mode: int = 0) -> np.ndarray:
    """Owner rank of every colour — the same deterministic rule as the plan (csrc/fdjac_abi.cu finish_colored_plan):
    mode 0 round-robin; mode 1 LPT (heaviest colour first onto the least-loaded rank, ties -> lowest rank)."""
    owner = np.zeros(n_colors, np.int32)
    if world <= 1:
        return owner
    if mode == 1:
        counts = np.asarray(counts, dtype=np.uint64)
        order = sorted(range(n_colors), key=lambda k: -int(counts[k]))   # stable: ties keep ascending colour
        load = [0] * world
        for k in order:
            best = min(range(world), key=lambda r: (load[r], r))
            owner[k] = best
            load[best] += int(counts[k]) + 1
    else:
        owner[:] = np.arange(n_colors) % world
    return owner


def entry_colors_csc(colptr, colorvec) -> np.ndarray:
    """0-based colour of every CSC entry (colour of its column); -1 where colorvec < 1."""
    colptr = np.asarray(colptr, dtype=np.int64)
    cv = np.asarray(colorvec, dtype=np.int64)
    per_col = np.diff(colptr)
    ec = np.repeat(cv - 1, per_col)
    ec[ec < 0] = -1
    return ec


class GatherPlan:
    """Index lists for the pack / all_gather / unpack exchange of owned Jacobian entries."""

    def __init__(self, entry_color: np.ndarray, owner: np.ndarray, world: int, device):
        self.world = world
        own = np.where(entry_color >= 0, owner[np.clip(entry_color, 0, max(len(owner) - 1, 0))] if len(owner) else 0, 0)
        self.idx: List[torch.Tensor] = []
        for r in range(world):
            self.idx.append(torch.from_numpy(np.nonzero(own == r)[0].astype(np.int64)).to(device))
        self.pad = max((int(i.numel()) for i in self.idx), default=0)

    def pack(self, values: torch.Tensor, rank: int) -> torch.Tensor:
        out = torch.zeros(self.pad, dtype=values.dtype, device=values.device)
        out[: self.idx[rank].numel()] = values.index_select(0, self.idx[rank])
        return out

    def unpack(self, values: torch.Tensor, gathered: List[torch.Tensor], skip_rank: Optional[int] = None):
        for r, g in enumerate(gathered):
            if r == skip_rank:
                continue
            values.index_copy_(0, self.idx[r], g[: self.idx[r].numel()])


def allgather_owned(values: torch.Tensor, gp: GatherPlan, rank: int, group=None):
    """values holds this rank's owned entries in place; on return it holds everybody's (gloo or nccl)."""
    mine = gp.pack(values, rank)
    gathered = [torch.empty_like(mine) for _ in range(gp.world)]
    dist.all_gather(gathered, mine, group=group)
    gp.unpack(values, gathered, skip_rank=rank)
    return values


# ------------------------------------------------------------------------------------------------ device side
class IpcBuffer:
    """A float64 device buffer from a dedicated cudaMalloc (so its CUDA IPC handle maps the buffer itself)."""

    def __init__(self, count: int, device: torch.device):
        self.count = int(count)
        p = C.c_void_p()
        with torch.cuda.device(device):
            L.check(L.lib().fdb_device_alloc(C.byref(p), max(self.count, 2) * 8))
        self.ptr = p.value
        self.device = device
        self.tensor = torch.as_tensor(api._DevArray(self.ptr, (self.count,)), device=device)

    def handle(self) -> bytes:
        h = C.create_string_buffer(64)
        L.check(L.lib().fdb_ipc_get_handle(C.c_void_p(self.ptr), h))
        return h.raw

    def free(self):
        if self.ptr:
            with torch.cuda.device(self.device):
                L.lib().fdb_device_free(C.c_void_p(self.ptr))
            self.ptr = 0


class DeviceBarrier:
    """fdb_sync: the device-side barrier that orders one-process-per-GPU ranks without NCCL on the data path — a flag
    block per rank in peer memory (CUDA IPC), one tiny kernel per barrier enqueued on the caller's stream
    (st.release.sys to every peer, ld.acquire.sys until every peer has signalled).  torch.distributed is used once, at
    construction, to exchange the 64-byte IPC handles."""

    def __init__(self, device, group=None):
        self.device = torch.device(device)
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(L.lib().fdb_sync_create(C.byref(h), self.rank, self.world, api._device_index(self.device)))
        self._h = h
        flags = C.c_void_p()
        L.check(L.lib().fdb_sync_flags(self._h, C.byref(flags)))
        mine = C.create_string_buffer(64)
        L.check(L.lib().fdb_ipc_get_handle(flags, mine))
        handles = [None] * self.world
        dist.all_gather_object(handles, mine.raw, group=group)
        self._mapped = []
        ptrs = (C.c_void_p * self.world)()
        with torch.cuda.device(self.device):
            for r, hb in enumerate(handles):
                if r == self.rank:
                    ptrs[r] = flags.value
                    continue
                p = C.c_void_p()
                L.check(L.lib().fdb_ipc_open(hb, C.byref(p)))
                self._mapped.append(p.value)
                ptrs[r] = p.value
            L.check(L.lib().fdb_sync_set_peers(self._h, ptrs))
        dist.barrier(group=group)          # every rank has mapped every flag block before the first signal

    def wait(self, stream=None):
        """Enqueue one barrier on `stream` (default: the current stream of the device)."""
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        L.check(L.lib().fdb_sync_barrier(self._h, C.c_void_p(stream)))

    def close(self):
        if self._h is None:
            return
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            for p in self._mapped:
                L.lib().fdb_ipc_close(C.c_void_p(p))
            self._mapped = []
        try:
            dist.barrier(group=self.group)       # nobody frees a flag block a peer may still signal
        except Exception:
            pass
        L.lib().fdb_sync_destroy(self._h)
        self._h = None


class GroupJacobian:
    """ONE process driving several GPUs through fdb_group_* (the C ABI's own multi-GPU entry: no torch.distributed, no
    NCCL): colours (dense: column blocks) are partitioned over `devices`; devices[0] is the root and owns x, J and fx;
    every member's scatter kernel stores the entries it owns straight into the root's J over NVLink.  Device ordinals
    may repeat ([0, 0] runs two members on one GPU — the same code path on a single-GPU box).

    fs: one f! per member (NativeFn whose context lives on that member's device, or a Python callable)."""

    def __init__(self, J, colorvec, fdtype, devices, *, sparsity=api._DEFAULT, **plan_kw):
        self.devices = [api._device_index(d) if not isinstance(d, int) else int(d) for d in devices]
        self.root = torch.device("cuda", self.devices[0])
        self.J, self.fdtype = J, fdtype
        if sparsity is api._DEFAULT:
            sparsity = J if api._has_sparsestruct(J) else None
        fd = api._fdtype_code(fdtype)
        o = api._opts(fd, self.devices[0], **plan_kw)
        n_dev = len(self.devices)
        devs = (C.c_int * n_dev)(*self.devices)
        h = C.c_void_p()
        cv_ptr, k0 = api._index_ptr(colorvec)
        self._keep = [k0]
        lib = L.lib()
        if sparsity is None:
            m, n, ld = api._dense_ld(J)
            L.check(lib.fdb_group_create_dense(C.byref(h), n_dev, devs, m, n, ld, C.byref(o)))
        elif isinstance(sparsity, api.SparseMatrixCSC):
            cp, k1 = api._index_ptr(sparsity.colptr)
            rv, k2 = api._index_ptr(sparsity.rowval)
            self._keep += [k1, k2]
            if isinstance(J, api.SparseMatrixCSC):
                same = J is sparsity or (J.colptr is sparsity.colptr and J.rowval is sparsity.rowval)
                jcp = jrv = None
                if not same:
                    jcp, k3 = api._index_ptr(J.colptr)
                    jrv, k4 = api._index_ptr(J.rowval)
                    self._keep += [k3, k4]
                L.check(lib.fdb_group_create_csc(C.byref(h), n_dev, devs, sparsity.m, sparsity.n, cp, rv, L.FDB_J_CSC_NZVAL,
                                                 jcp, jrv, 0, cv_ptr, C.byref(o)))
            else:
                m, n, ld = api._dense_ld(J)
                L.check(lib.fdb_group_create_csc(C.byref(h), n_dev, devs, m, n, cp, rv, L.FDB_J_DENSE, None, None, ld,
                                                 cv_ptr, C.byref(o)))
        elif isinstance(sparsity, api.BandedMatrix):
            if isinstance(J, api.BandedMatrix):
                L.check(lib.fdb_group_create_banded(C.byref(h), n_dev, devs, sparsity.m, sparsity.n, sparsity.l, sparsity.u,
                                                    L.FDB_J_BAND, 0, cv_ptr, C.byref(o)))
            else:
                m, n, ld = api._dense_ld(J)
                L.check(lib.fdb_group_create_banded(C.byref(h), n_dev, devs, m, n, sparsity.l, sparsity.u, L.FDB_J_DENSE, ld,
                                                    cv_ptr, C.byref(o)))
        else:
            raise TypeError(f"GroupJacobian: unsupported sparsity type {type(sparsity)}")
        self._h = h
        self._fin = api.weakref.finalize(self, lib.fdb_group_destroy, C.c_void_p(h.value))
        self.plans = []
        for i in range(n_dev):
            p = C.c_void_p()
            L.check(lib.fdb_group_plan(self._h, i, C.byref(p)))
            self.plans.append(api.Plan(p.value, owned=False))
        self._fns = None

    def run(self, fs, x, fx=None, f_in=None, *, relstep=None, absstep=None, dir=True, stream=None):
        m, n = api._shape_of(self.J)
        if len(fs) != len(self.devices):
            raise ValueError("one f! per member")
        if self._fns is None or self._fns[0] is not fs:
            if all(isinstance(f, api.NativeFn) for f in fs):
                wrapped = [api._as_fn(f, m, n, torch.device("cuda", d), 1) for f, d in zip(fs, self.devices)]
            else:
                # Python callables: ONE trampoline (the ABI takes one f and a context per member); the context is the
                # member index + 1, the trampoline forwards to that member's wrapper (tensors on that member's device)
                per = [api._PyFn(f, m, n, torch.device("cuda", d), bool(getattr(f, "batched", False))) for f, d in zip(fs, self.devices)]

                def tramp(ctx, p_fx, p_x, batch, ldfx, ldx, stream, _per=per):
                    return _per[int(ctx or 1) - 1]._tramp(None, p_fx, p_x, batch, ldfx, ldx, stream)

                cf = L.FDB_FN(tramp)
                wrapped = [(L.fn_address(cf), C.c_void_p(i + 1), per[i]) for i in range(len(per))]
                self._tramp_keep = cf
            self._fns = (fs, wrapped)
        wrapped = self._fns[1]
        addr = wrapped[0][0]
        if any(w[0] != addr for w in wrapped):
            raise ValueError("all members must share one f! entry point (contexts may differ per device)")
        ctxs = (C.c_void_p * len(wrapped))(*[C.cast(w[1], C.c_void_p).value if w[1] is not None else None for w in wrapped])
        if stream is None:
            stream = torch.cuda.current_stream(self.root).cuda_stream
        jv = api._j_values(self.J)
        with torch.cuda.device(self.root):
            st = L.lib().fdb_group_jacobian(self._h, addr, ctxs, x.data_ptr(), jv.data_ptr(),
                                            None if fx is None else fx.data_ptr(), None if f_in is None else f_in.data_ptr(),
                                            L.STEP_DEFAULT if relstep is None else float(relstep),
                                            L.STEP_DEFAULT if absstep is None else float(absstep), float(dir), C.c_void_p(stream))
        for w in wrapped:
            if st == L.FDB_ERR_CALLBACK and w[2] is not None and w[2].exc is not None:
                exc, w[2].exc = w[2].exc, None
                raise exc
        L.check(st)

    def synchronize(self):
        for d in set(self.devices):
            torch.cuda.synchronize(d)

    def close(self):
        self.synchronize()
        self._fin()


class ShardedJacobian:
    """finite_difference_jacobian! with the colours of `cache.colorvec` sharded over the ranks of the default process
    group, one process per GPU (the cache must have been built with rank=, world=).

    mode="p2p" (default): no collective on the data path.  Rank 0's nzval lives in a CUDA-IPC buffer that every rank
    maps; the plans are created with shared_j, so each rank's diff+scatter kernel stores the entries it owns STRAIGHT
    into rank 0's nzval over NVLink (the literal "final gather of Jacobian columns", fused into the kernel), and the
    ranks are ordered by the device-side flag barrier of the C ABI (fdb_sync: DeviceBarrier) — no NCCL call per Jacobian.
      gather="root"     rank 0 ends with the complete J;
      gather="all"      + one NCCL broadcast of nzval (bulk, full NVLink bandwidth): every rank ends with it;
      gather="all_p2p"  every value is stored to EVERY peer by the scatter kernel (fdb_plan_set_peers).  Fine up to 4
                        GPUs, 16 ms at 8 (7 peer apertures x scattered 64-byte runs) — kept for experiments only.
    mode="nccl": pack owned entries -> all_gather -> unpack (fallback; also what the CPU/gloo tests exercise)."""

    def __init__(self, J: api.SparseMatrixCSC, cache: api.JacobianCache, n: int, device, mode: str = "p2p", group=None,
                 pre_sync: bool = True, gather: str = "all", barrier: str = "device"):
        if gather not in ("all", "root", "all_p2p"):
            raise ValueError("gather must be 'all', 'root' or 'all_p2p'")
        if barrier not in ("device", "nccl"):
            raise ValueError("barrier must be 'device' (fdb_sync) or 'nccl' (all_reduce of a flag word)")
        self.gather = gather
        if not isinstance(J, api.SparseMatrixCSC):
            raise TypeError("ShardedJacobian shards CSC Jacobians (dense / banded plans: GroupJacobian, or rank/world plans "
                            "with their own column slabs)")
        self.J, self.cache, self.n, self.device, self.group = J, cache, n, torch.device(device), group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.pre_sync = pre_sync
        self.shared = mode == "p2p" and gather in ("root", "all")
        if self.shared:
            cache._plan_kw["shared_j"] = True
        self.plan = cache.plan_for(J, cache.sparsity, cache.colorvec, n)
        self.flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.mode = mode
        self._peers = []
        self._root_ptr = None
        self._J_run = J
        self._ipc: Optional[IpcBuffer] = None
        self.barrier: Optional[DeviceBarrier] = None
        self._barrier_kind = barrier
        if mode == "p2p":
            try:
                self._setup_p2p()
            except Exception as e:  # no IPC / no peer access: fall back to the NCCL exchange
                self.mode = "nccl"
                self._fallback_reason = repr(e)
        if self.mode == "nccl":
            cp = J.colptr.cpu().numpy() if isinstance(J.colptr, torch.Tensor) else np.asarray(J.colptr)
            cv = cache.colorvec
            cv = cv.cpu().numpy() if isinstance(cv, torch.Tensor) else np.asarray(list(cv) if isinstance(cv, range) else cv)
            self.gp = GatherPlan(entry_colors_csc(cp, cv), self.plan.color_owner(), self.world, self.device)

    def _setup_p2p(self):
        nnz = self.J.nzval.numel()
        handles = [None] * self.world
        if self.shared:
            mine = None
            if self.rank == 0:
                self._ipc = IpcBuffer(nnz, self.device)
                self._ipc.tensor.copy_(self.J.nzval)
                self.J.nzval = self._ipc.tensor             # rank 0's values now live in the IPC-exportable buffer
                mine = self._ipc.handle()
            dist.all_gather_object(handles, mine, group=self.group)
            if self.rank != 0:
                p = C.c_void_p()
                with torch.cuda.device(self.device):
                    L.check(L.lib().fdb_ipc_open(handles[0], C.byref(p)))
                self._root_ptr = p.value
                # rank 0's nzval as mapped into this process: a raw pointer (api.PeerValues), never a torch tensor
                self._J_run = api.SparseMatrixCSC(self.J.m, self.J.n, self.J.colptr, self.J.rowval,
                                                  api.PeerValues(self._root_ptr, nnz))
        else:
            self._ipc = IpcBuffer(nnz, self.device)
            self._ipc.tensor.copy_(self.J.nzval)
            self.J.nzval = self._ipc.tensor
            dist.all_gather_object(handles, self._ipc.handle(), group=self.group)
            ptrs = []
            with torch.cuda.device(self.device):
                for r, h in enumerate(handles):
                    if r == self.rank:
                        continue
                    p = C.c_void_p()
                    L.check(L.lib().fdb_ipc_open(h, C.byref(p)))
                    ptrs.append(p.value)
            self._peers = ptrs
            self.plan.set_peers(ptrs)
        if self._barrier_kind == "device":
            self.barrier = DeviceBarrier(self.device, self.group)
        ok = torch.ones(1, device=self.device)
        dist.all_reduce(ok, group=self.group)

    def _sync(self):
        if self.barrier is not None:
            self.barrier.wait()
        else:
            dist.all_reduce(self.flag, group=self.group)

    def run(self, f, x, **kw):
        if self.mode == "p2p" and self.pre_sync:
            self._sync()                                   # nobody may still be reading J (stream-ordered)
        api.finite_difference_jacobian_(self._J_run, f, x, self.cache, **kw)
        if self.mode == "p2p":
            self._sync()                                   # every rank's scatter (incl. its NVLink stores) has completed
            if self.gather == "all":
                dist.broadcast(self.J.nzval, src=0, group=self.group)
        else:
            allgather_owned(self.J.nzval, self.gp, self.rank, self.group)
        return None

    def close(self):
        torch.cuda.synchronize(self.device)
        if self._peers:
            self.plan.set_peers([])
            with torch.cuda.device(self.device):
                for p in self._peers:
                    L.lib().fdb_ipc_close(C.c_void_p(p))
            self._peers = []
        if self._root_ptr:
            with torch.cuda.device(self.device):
                L.lib().fdb_ipc_close(C.c_void_p(self._root_ptr))
            self._root_ptr = None
            self._J_run = self.J
        if self.barrier is not None:
            self.barrier.close()
            self.barrier = None
        if self.group is None or dist.is_initialized():
            try:
                dist.barrier(group=self.group)
            except Exception:
                pass
        if self._ipc is not None:
            self._ipc.free()
            self._ipc = None


# ------------------------------------------------------------------------------------------------ column-block shards
# problems with fewer colours than GPUs (C2: 3 colours, C3: 5) cannot be spread by colour.  A
# contiguous block of COLUMNS can: the block's entries live in a row range [r0, r1), those rows depend on an x range
# [x0, x1) (the block plus a halo), and everything the colour loop does for the block happens inside those ranges —
# with a slice-aware f! (rows [r0, r1) from x[x0:x1]) there is no exchange at all on the data path.  The only global
# quantity is the step size of each colour (norm over ALL colour-k components of x, jacobians.jl:559-561): every rank
# holds the full x and runs the K2 pass on it (fdb_color_eps), the shard plan takes the result as external step sizes.
# With the same step sizes a shard's values are bit-identical to its segment of the unsharded nzval.
def column_blocks(colptr, world: int) -> List[int]:
    """Block boundaries b[0..world] (0-based columns) balancing the number of stored entries per block."""
    cp = colptr.cpu().numpy() if isinstance(colptr, torch.Tensor) else np.asarray(colptr)
    n = len(cp) - 1
    nnz = int(cp[-1] - 1)
    bounds = [0]
    for r in range(1, world):
        target = 1 + (nnz * r) // world
        bounds.append(int(min(max(np.searchsorted(cp, target, side="left"), bounds[-1]), n)))
    bounds.append(n)
    return bounds


def block_geometry(colptr: torch.Tensor, rowval: torch.Tensor, c0: int, c1: int) -> dict:
    """Everything a column block [c0, c1) of a CSC pattern (Int64, 1-based, rows sorted inside a column) needs:
    p0, p1   0-based slot range of the block in rowval / nzval,
    r0, r1   the row range its entries touch (0-based, half-open),
    x0, x1   hull of the columns with an entry in those rows = the x slice a slice-aware f! must see (contains [c0, c1)),
    colptr_loc / rowval_loc   the block's pattern over the columns [x0, x1) and rows [r0, r1), 1-based: only the owned
             columns carry entries."""
    cp, rv = colptr, rowval
    p0, p1 = int(cp[c0]) - 1, int(cp[c1]) - 1
    seg = rv[p0:p1]
    if seg.numel() == 0:
        r0 = r1 = 0
        x0, x1 = c0, c1
    else:
        r0, r1 = int(seg.min()) - 1, int(seg.max())
        cnt = cp[1:] - cp[:-1]
        first = rv[(cp[:-1] - 1).clamp(max=max(rv.numel() - 1, 0))]
        last = rv[(cp[1:] - 2).clamp(min=0)]
        touch = (cnt > 0) & (last - 1 >= r0) & (first - 1 < r1)
        idx = torch.nonzero(touch).reshape(-1)
        x0, x1 = min(int(idx[0]), c0), max(int(idx[-1]) + 1, c1)
    j = torch.arange(x0, x1 + 1, device=cp.device)
    colptr_loc = (cp[j.clamp(c0, c1)] - p0).contiguous()
    rowval_loc = (seg - r0).contiguous()
    return dict(p0=p0, p1=p1, r0=r0, r1=r1, x0=x0, x1=x1, colptr_loc=colptr_loc, rowval_loc=rowval_loc)


class EpsPlan:
    """Step sizes of every colour for a full-length x: the K2 pass alone (fdb_eps_plan_create / fdb_color_eps)."""

    def __init__(self, n: int, colorvec, fdtype, device):
        self.device = torch.device(device)
        h = C.c_void_p()
        o = api._opts(api._fdtype_code(fdtype), api._device_index(device))
        cv_ptr, keep = api._index_ptr(colorvec)
        with torch.cuda.device(self.device):
            L.check(L.lib().fdb_eps_plan_create(C.byref(h), int(n), cv_ptr, C.byref(o)))
        self.plan = api.Plan(h.value, keep=(keep,))
        self.n_colors = self.plan.info()["n_colors"]
        self.eps = torch.zeros(max(self.n_colors, 1), dtype=torch.float64, device=self.device)

    def compute(self, x: torch.Tensor, relstep=None, absstep=None, dir=True, stream=None) -> torch.Tensor:
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            L.check(L.lib().fdb_color_eps(self.plan.handle, x.data_ptr(), L.STEP_DEFAULT if relstep is None else float(relstep),
                                          L.STEP_DEFAULT if absstep is None else float(absstep), float(dir), self.eps.data_ptr(),
                                          C.c_void_p(stream)))
        return self.eps


class ColumnBlockJacobian:
    """The colour loop of finite_difference_jacobian! for the columns [c0, c1) of a CSC Jacobian.

    f_factory(row0, row1, x0, x1) must return the slice-aware f! (NativeFn or Python callable f(fx, x)) that computes
    rows [row0, row1) of f from x[x0:x1] (0-based, half-open).  run(x, eps) writes the block's values into
    values_ptr (default: J.nzval) at slots [p0, p1) — the same positions the unsharded call writes."""

    def __init__(self, J: api.SparseMatrixCSC, colorvec, fdtype, c0: int, c1: int, device, f_factory, *, max_batch=1,
                 use_graph=False, no_drift=False):
        self.device = torch.device(device)
        self.fdtype = fdtype
        cp = J.colptr if isinstance(J.colptr, torch.Tensor) else torch.as_tensor(np.asarray(J.colptr))
        rv = J.rowval if isinstance(J.rowval, torch.Tensor) else torch.as_tensor(np.asarray(J.rowval))
        cp, rv = cp.to(self.device), rv.to(self.device)
        self.c0, self.c1 = int(c0), int(c1)
        g = block_geometry(cp, rv, self.c0, self.c1)
        self.p0, self.p1, self.r0, self.r1, self.x0, self.x1 = g["p0"], g["p1"], g["r0"], g["r1"], g["x0"], g["x1"]
        self.m_loc, self.n_loc = self.r1 - self.r0, self.x1 - self.x0
        colptr_loc, rowval_loc = g["colptr_loc"], g["rowval_loc"]
        if isinstance(colorvec, range):
            colorvec = np.arange(colorvec.start, colorvec.stop, colorvec.step, dtype=np.int64)
        cv = colorvec if isinstance(colorvec, torch.Tensor) else torch.as_tensor(np.asarray(colorvec, dtype=np.int64))
        self.cv_loc = cv[self.x0:self.x1].to(self.device).contiguous()
        self.sub = api.SparseMatrixCSC(self.m_loc, self.n_loc, colptr_loc, rowval_loc, None)
        self.values_default = J.nzval
        self.plan = api.make_plan(self.sub, self.sub, self.cv_loc, fdtype, self.n_loc, self.device, max_batch=max_batch,
                                  use_graph=use_graph, no_drift=no_drift)
        self.f = f_factory(self.r0, self.r1, self.x0, self.x1)
        self._fn = api._as_fn(self.f, self.m_loc, self.n_loc, self.device, 1)
        self._eps_set = None

    def run(self, x: torch.Tensor, eps: torch.Tensor, values_ptr: Optional[int] = None, dir=True, stream=None):
        addr, ctx, pyfn = self._fn
        if self._eps_set != eps.data_ptr():
            L.check(L.lib().fdb_plan_set_external_eps(self.plan.handle, eps.data_ptr()))
            self._eps_set = eps.data_ptr()
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        base = self.values_default.data_ptr() if values_ptr is None else int(values_ptr)
        with torch.cuda.device(self.device):
            st = L.lib().fdb_jacobian(self.plan.handle, addr, ctx, x.data_ptr() + 8 * self.x0, base + 8 * self.p0, None,
                                      None, L.STEP_DEFAULT, L.STEP_DEFAULT, float(dir), C.c_void_p(stream))
        if st == L.FDB_ERR_CALLBACK and pyfn is not None and pyfn.exc is not None:
            exc, pyfn.exc = pyfn.exc, None
            raise exc
        L.check(st)


class ColumnShardedJacobian:
    """One process per GPU, every rank owns a contiguous block of columns (entry-balanced).  Every rank needs the full
    x; J stays column-sharded (gather=None: rank r's J.nzval holds its slots [p0, p1)) or is assembled on rank 0
    (gather="root": the scatter kernel stores straight into rank 0's nzval over NVLink — contiguous 16-byte stores)."""

    def __init__(self, J: api.SparseMatrixCSC, colorvec, fdtype, device, f_factory, *, gather: Optional[str] = None,
                 group=None, max_batch=1, use_graph=False):
        if gather not in (None, "root"):
            raise ValueError("gather must be None or 'root'")
        self.J, self.device, self.group, self.gather = J, torch.device(device), group, gather
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.bounds = column_blocks(J.colptr, self.world)
        self.eps_plan = EpsPlan(J.n, colorvec, fdtype, device)
        self.block = ColumnBlockJacobian(J, colorvec, fdtype, self.bounds[self.rank], self.bounds[self.rank + 1], device,
                                         f_factory, max_batch=max_batch, use_graph=use_graph)
        self.flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._ipc: Optional[IpcBuffer] = None
        self._root_ptr = None
        self.barrier: Optional[DeviceBarrier] = None
        if gather == "root":
            self.barrier = DeviceBarrier(self.device, group)
            handles = [None] * self.world
            mine = None
            if self.rank == 0:
                self._ipc = IpcBuffer(J.nzval.numel(), self.device)
                self._ipc.tensor.copy_(J.nzval)
                J.nzval = self._ipc.tensor
                self.block.values_default = J.nzval
                mine = self._ipc.handle()
            dist.all_gather_object(handles, mine, group=group)
            if self.rank != 0:
                p = C.c_void_p()
                with torch.cuda.device(self.device):
                    L.check(L.lib().fdb_ipc_open(handles[0], C.byref(p)))
                self._root_ptr = p.value

    def run(self, x: torch.Tensor, dir=True):
        eps = self.eps_plan.compute(x, dir=dir)
        if self.gather == "root":
            self.barrier.wait()                               # rank 0 is done reading the previous J
        self.block.run(x, eps, values_ptr=self._root_ptr, dir=dir)
        if self.gather == "root":
            self.barrier.wait()                               # every block (incl. its NVLink stores) has completed

    def close(self):
        torch.cuda.synchronize(self.device)
        if self._root_ptr:
            with torch.cuda.device(self.device):
                L.lib().fdb_ipc_close(C.c_void_p(self._root_ptr))
            self._root_ptr = None
        if self.barrier is not None:
            self.barrier.close()
            self.barrier = None
        try:
            dist.barrier(group=self.group)
        except Exception:
            pass
        if self._ipc is not None:
            self._ipc.free()
