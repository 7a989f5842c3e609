"""One process per GPU: the colour set of a Jacobian is partitioned over the ranks.

Colours are independent units given x (each colour = perturb -> f! -> diff -> scatter into a disjoint set of J slots),
so there is NO collective on the data path until the end, where every rank needs the entries the others computed.
Two implementations of that final exchange:

  "p2p"  (default on NVLink/NVSwitch)  the diff+scatter kernel itself stores every value it owns into rank 0's nzval
         buffer (gather="root"; peer pointer from CUDA IPC, passed to the plan with fdb_plan_set_peers) — compute and
         gather are one kernel, overlapped with the next colour's f! on a side stream; a stream-ordered NCCL all-reduce
         of one flag word is the only barrier.  gather="all" adds one NCCL broadcast of nzval; gather="all_p2p" stores
         to every peer from the kernel (fine up to 4 GPUs, pathological at 8 — see ShardedJacobian).
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


class ShardedJacobian:
    """finite_difference_jacobian! with the colours of `cache.colorvec` sharded over the ranks of the default process
    group (the cache must have been built with rank=, world=).  After run() every rank holds the complete J."""

    def __init__(self, J: api.SparseMatrixCSC, cache: api.JacobianCache, n: int, device, mode: str = "p2p", group=None,
                 pre_sync: bool = True, gather: str = "all"):
        """gather="root": rank 0 ends with the complete J — every rank's scatter kernel stores its values straight into
        rank 0's nzval (the literal "final gather of Jacobian columns");
        gather="all": every rank does: the same fused gather to rank 0, then one NCCL broadcast of nzval (bulk, full
        NVLink bandwidth);
        gather="all_p2p": every value is stored to EVERY peer by the scatter kernel.  Fine up to 4 GPUs (3.35x at 4), but
        measured 16 ms at 8 GPUs (7 peer apertures x scattered 64-byte runs) — kept for experiments only."""
        if gather not in ("all", "root", "all_p2p"):
            raise ValueError("gather must be 'all', 'root' or 'all_p2p'")
        self.gather = gather
        if not isinstance(J, api.SparseMatrixCSC):
            raise TypeError("ShardedJacobian shards CSC Jacobians (dense plans shard columns: Plan.dense_range())")
        self.J, self.cache, self.n, self.device, self.group = J, cache, n, torch.device(device), group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.pre_sync = pre_sync
        self.plan = cache.plan_for(J, cache.sparsity, cache.colorvec, n)
        self.flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.mode = mode
        self._peers = []
        self._ipc: Optional[IpcBuffer] = None
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
        self._ipc = IpcBuffer(nnz, self.device)
        self._ipc.tensor.copy_(self.J.nzval)
        self.J.nzval = self._ipc.tensor                 # J's values now live in the IPC-exportable buffer
        handles = [None] * self.world
        dist.all_gather_object(handles, self._ipc.handle(), group=self.group)
        ptrs = []
        with torch.cuda.device(self.device):
            for r, h in enumerate(handles):
                if r == self.rank or (self.gather != "all_p2p" and (r != 0 or self.rank == 0)):
                    continue
                p = C.c_void_p()
                L.check(L.lib().fdb_ipc_open(h, C.byref(p)))
                ptrs.append(p.value)
        self._peers = ptrs
        self.plan.set_peers(ptrs)
        ok = torch.ones(1, device=self.device)
        dist.all_reduce(ok, group=self.group)

    def run(self, f, x, **kw):
        if self.mode == "p2p" and self.pre_sync:
            dist.all_reduce(self.flag, group=self.group)    # nobody may still be reading J (stream-ordered)
        api.finite_difference_jacobian_(self.J, f, x, self.cache, **kw)
        if self.mode == "p2p":
            dist.all_reduce(self.flag, group=self.group)    # every rank's scatter (incl. its peer stores) has completed
            if self.gather == "all":
                dist.broadcast(self.J.nzval, src=0, group=self.group)
        else:
            allgather_owned(self.J.nzval, self.gp, self.rank, self.group)
        return None

    def close(self):
        if self._peers:
            self.plan.set_peers([])
            with torch.cuda.device(self.device):
                for p in self._peers:
                    L.lib().fdb_ipc_close(C.c_void_p(p))
            self._peers = []
        torch.cuda.synchronize(self.device)
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
            L.check(L.lib().fdb_color_eps(self.plan.handle, x.data_ptr(), 0.0 if relstep is None else float(relstep),
                                          0.0 if absstep is None else float(absstep), float(dir), self.eps.data_ptr(),
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
                                      None, 0.0, 0.0, float(dir), C.c_void_p(stream))
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
        if gather == "root":
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
            dist.all_reduce(self.flag, group=self.group)      # rank 0 is done reading the previous J
        self.block.run(x, eps, values_ptr=self._root_ptr, dir=dir)
        if self.gather == "root":
            dist.all_reduce(self.flag, group=self.group)      # every block (incl. its peer stores) has completed

    def close(self):
        torch.cuda.synchronize(self.device)
        if self._root_ptr:
            with torch.cuda.device(self.device):
                L.lib().fdb_ipc_close(C.c_void_p(self._root_ptr))
            self._root_ptr = None
        try:
            dist.barrier(group=self.group)
        except Exception:
            pass
        if self._ipc is not None:
            self._ipc.free()
