"""Device-resident graph + the walk launcher: the host side of K0/K1/K2.

``DeviceGraph`` replaces the reference's ``df_adj`` frame of pickled adjacency strings
(fugue.py:130, randomwalk.py:266-275): a packed CSR in HBM with every vertex's
first-order alias table folded into its arc records (include/n2v_b200.h).
PyTorch tensors own all device memory; kernels run on torch's current stream.
"""
import ctypes as C
import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib


def _as_device(x, device, np_dtype, torch_dtype) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch_dtype).contiguous()
    arr = np.ascontiguousarray(x).astype(np_dtype, copy=False)
    if not arr.flags.writeable:          # pandas >= 3 hands out read-only views; torch wants writable memory
        arr = arr.copy()
    return torch.as_tensor(arr, device=device)


def _as_device_i32(x, device) -> torch.Tensor:
    """Vertex ids as int32 on the device.  Wider integer inputs are range-checked first: a silent
    wrap of an id >= 2^31 would alias a valid vertex (ids outside [0, V) are rejected by K0)."""
    if isinstance(x, torch.Tensor):
        if x.dtype in (torch.int64, torch.uint8, torch.int16) and x.numel():
            lo, hi = int(x.min()), int(x.max())
            if lo < -(1 << 31) or hi >= (1 << 31):
                raise ValueError(f"vertex ids must fit int32, got range [{lo}, {hi}]")
    else:
        arr = np.asarray(x)
        if arr.dtype.kind in "iu" and arr.dtype.itemsize > 4 and arr.size:
            lo, hi = int(arr.min()), int(arr.max())
            if lo < -(1 << 31) or hi >= (1 << 31):
                raise ValueError(f"vertex ids must fit int32, got range [{lo}, {hi}]")
    return _as_device(x, device, np.int32, torch.int32)


def _as_device_f64(x, device) -> torch.Tensor:
    return _as_device(x, device, np.float64, torch.float64)


class DeviceGraph:
    """Sorted CSR + alias records in HBM.

    Layout (one replica):
      vtx   int32[V, 4]   16 B/vertex  {base u32, deg u32, hbase u32, wsum f32}
      arcs  int32[A, 8]   32 B/arc     {thr, dst, alias_dst, alias_idx, dst_base, dst_deg, adst_base, adst_deg}
                                       = one sector per walk trial, landing vertex's header included
      hash  int32[B, 8]   ~8 B/arc     per-vertex neighbour hash sets, 32 B buckets (membership test)
      col   int32[A]       4 B/arc     neighbour ids, ascending per vertex (exact fallback, parity)
      weight f64[A]        8 B/arc     reference weights (exact fallback, parity outputs)
    ``alias`` / ``probs`` (the reference's tables, bit-exact) are kept only on request.
    """

    def __init__(self):
        self.n_vertices = 0
        self.n_arcs = 0
        self.flags = 0
        self.vtx = self.arcs = self.col = self.weight = self.hash = None
        self.alias = self.probs = self.perm = self.ratio = None
        self.device = None
        self.sum_mode = "naive"
        self._struct = None

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_arcs(cls, src, dst, weight=None, n_vertices: Optional[int] = None, device=None,
                  sum_mode: str = "naive", keep_tables: bool = False, keep_perm: bool = False) -> "DeviceGraph":
        """K0 + K1.  src/dst: int ids (tensor or array); weight: fp64 or None (=1.0)."""
        _lib.require_cuda()
        lib = _lib.load()
        if sum_mode not in _lib.SUM_MODE:
            raise ValueError(f"unknown sum_mode {sum_mode!r}")
        device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        g = cls()
        g.device, g.sum_mode = device, sum_mode
        with torch.cuda.device(device):
            s = _as_device_i32(src, device)
            d = _as_device_i32(dst, device)
            w = None if weight is None else _as_device_f64(weight, device)
            if s.numel() != d.numel() or (w is not None and w.numel() != s.numel()):
                raise ValueError("src, dst and weight must have the same length")
            n_arcs = int(s.numel())
            if n_vertices is None:
                n_vertices = int(torch.maximum(s.max(), d.max())) + 1 if n_arcs else 0      # one host round trip
            g.n_vertices, g.n_arcs = int(n_vertices), n_arcs
            stream = _lib.current_stream_ptr()
            g.vtx = torch.empty((g.n_vertices, 4), dtype=torch.int32, device=device)
            g.col = torch.empty(n_arcs, dtype=torch.int32, device=device)
            g.weight = torch.empty(n_arcs, dtype=torch.float64, device=device)
            if keep_perm:
                g.perm = torch.empty(n_arcs, dtype=torch.int64, device=device)
            nbytes = int(lib.n2v_csr_scratch_bytes(n_arcs, g.n_vertices))
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=device)
            flags = C.c_uint32(0)
            _lib.check(lib.n2v_csr_build(_lib.ptr(s), _lib.ptr(d), _lib.ptr(w), n_arcs, g.n_vertices, g.n_vertices,
                                         _lib.ptr(g.vtx), _lib.ptr(g.col), _lib.ptr(g.weight), _lib.ptr(g.perm),
                                         _lib.ptr(scratch), nbytes, C.byref(flags), stream), "n2v_csr_build")
            g.flags = int(flags.value)
            del scratch, s, d, w
            n_buckets = int(lib.n2v_hash_buckets_bound(n_arcs, g.n_vertices))
            g.hash = torch.empty((n_buckets, 8), dtype=torch.int32, device=device)
            _lib.check(lib.n2v_hash_build(_lib.ptr(g.vtx), _lib.ptr(g.col), g.n_vertices, n_arcs, _lib.ptr(g.hash),
                                          n_buckets, stream), "n2v_hash_build")
            g.arcs = torch.empty((n_arcs, 8), dtype=torch.int32, device=device)
            probs = torch.empty(n_arcs, dtype=torch.float64, device=device)
            alias = torch.empty(n_arcs, dtype=torch.int32, device=device) if keep_tables else None
            work = torch.empty(n_arcs, dtype=torch.int32, device=device)
            n_zero = C.c_int64(0)
            _lib.check(lib.n2v_alias_build(_lib.ptr(g.vtx), None, _lib.ptr(g.col), _lib.ptr(g.weight), g.n_vertices,
                                           n_arcs, _lib.SUM_MODE[sum_mode], _lib.ptr(alias), _lib.ptr(probs),
                                           _lib.ptr(g.arcs), _lib.ptr(work), C.byref(n_zero), stream),
                       "n2v_alias_build")
            if keep_tables:
                g.alias, g.probs = alias, probs
        g._make_struct()
        return g

    def _make_struct(self):
        st = _lib.Graph()
        st.n_vertices, st.n_arcs, st.flags = self.n_vertices, self.n_arcs, self.flags
        st.n_parts, st.part_size = 1, max(self.n_vertices, 1)
        st.parts[0].vtx = self.vtx.data_ptr()
        st.parts[0].arcs = self.arcs.data_ptr()
        st.parts[0].col = self.col.data_ptr()
        st.parts[0].weight = self.weight.data_ptr()
        st.parts[0].hash = self.hash.data_ptr()
        self._struct = st

    # ------------------------------------------------------------------ views
    @property
    def struct(self) -> "_lib.Graph":
        return self._struct

    def degrees(self) -> torch.Tensor:
        return self.vtx[:, 1].clone()

    def start_vertices(self) -> torch.Tensor:
        """Vertices with at least one out-arc, ascending -- ``walk_start = df_adj[["id"]]``
        (fugue.py:132): only they start walks."""
        return torch.nonzero(self.vtx[:, 1] != 0).view(-1).to(torch.int32)

    def nbytes(self) -> int:
        return sum(int(t.numel()) * t.element_size() for t in (self.vtx, self.arcs, self.col, self.weight, self.hash))

    def to_host(self) -> Dict[str, np.ndarray]:
        """Plain host arrays (for tests and the oracle's replay)."""
        vtx = self.vtx.cpu().numpy()
        arcs = self.arcs.cpu().numpy()
        out = {
            "base": vtx[:, 0].copy().view(np.uint32).astype(np.uint64),
            "deg": vtx[:, 1].copy().view(np.uint32),
            "hbase": vtx[:, 2].copy().view(np.uint32),
            "wsum": vtx[:, 3].copy().view(np.float32),
            "hash": self.hash.cpu().numpy(),
            "thr": arcs[:, 0].copy().view(np.uint32),
            "dst": arcs[:, 1].copy(),
            "alias_dst": arcs[:, 2].copy(),
            "alias_idx": arcs[:, 3].copy(),
            "dst_base": arcs[:, 4].copy().view(np.uint32), "dst_deg": arcs[:, 5].copy().view(np.uint32),
            "adst_base": arcs[:, 6].copy().view(np.uint32), "adst_deg": arcs[:, 7].copy().view(np.uint32),
            "col": self.col.cpu().numpy(),
            "weight": self.weight.cpu().numpy(),
        }
        if self.ratio is not None:
            out["ratio"] = self.ratio.cpu().numpy()
        if self.alias is not None:
            out["alias"] = self.alias.cpu().numpy()
            out["probs"] = self.probs.cpu().numpy()
        return out

    def ensure_ratio(self) -> None:
        """Build the per-arc {fwd, rev} return-mass ratios (general fold) once, on demand."""
        if self.ratio is not None or self.n_arcs == 0 or self._struct.n_parts != 1:
            return
        lib = _lib.load()
        with torch.cuda.device(self.device):
            self.ratio = torch.empty((self.n_arcs, 2), dtype=torch.float32, device=self.device)
            _lib.check(lib.n2v_ratio_build(C.byref(self._struct), _lib.ptr(self.ratio), _lib.current_stream_ptr()),
                       "n2v_ratio_build")
            self._struct.parts[0].ratio = self.ratio.data_ptr()

    # ------------------------------------------------------------------ K2
    def walk(self, start, num_walks: int, walk_length: int, return_param: float = 1.0,
             inout_param: float = 1.0, seed: Optional[int] = None, collect_stats: bool = False,
             out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, Optional[Dict[str, int]]]:
        """Launch the walk kernel.  Returns (walks[W, L+1] view of a pitch-padded buffer,
        alive[W] bool, stats or None).  Row w = start index * num_walks + walk number."""
        lib = _lib.load()
        _lib.tune_device(self.device.index)
        if return_param == 0 or inout_param == 0:
            raise ValueError(f"Zero return ({return_param}) or inout ({inout_param}) parameter!")
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        need = _lib.GRAPH_UNIT_WEIGHT | _lib.GRAPH_SYMMETRIC | _lib.GRAPH_SIMPLE
        if (self.flags & need) != need and 1.0 / return_param > max(1.0, 1.0 / inout_param):
            self.ensure_ratio()      # weighted / directed / multi-arc graph with small p: general fold
        with torch.cuda.device(self.device):
            start_t = _as_device_i32(start, self.device)
            n_start = int(start_t.numel())
            W = n_start * int(num_walks)
            pitch = (int(walk_length) + 1 + 7) // 8 * 8
            if out is None:
                out = torch.empty((W, pitch), dtype=torch.int32, device=self.device)
            elif out.shape != (W, pitch) or out.dtype != torch.int32 or not out.is_contiguous():
                raise ValueError(f"out must be a contiguous int32 tensor of shape {(W, pitch)}")
            alive = torch.empty(W, dtype=torch.uint8, device=self.device)
            stats = torch.zeros(8, dtype=torch.int64, device=self.device) if collect_stats else None
            _lib.check(lib.n2v_walk(C.byref(self._struct), _lib.ptr(start_t), n_start, int(num_walks),
                                    int(walk_length), float(return_param), float(inout_param),
                                    C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), _lib.ptr(out), pitch, _lib.ptr(alive),
                                    _lib.ptr(stats), _lib.current_stream_ptr()), "n2v_walk")
            st = None
            if collect_stats:
                st = dict(zip(_lib.WALK_STAT_NAMES, stats.cpu().tolist()))
        return out[:, : walk_length + 1], alive.bool(), st


def walk_to_host(graph, start, num_walks: int, walk_length: int, return_param: float, inout_param: float,
                 seed: Optional[int], out_host: torch.Tensor, chunk_walkers: Optional[int] = None):
    """Walk and deliver the rows into a PINNED host matrix, with the device->host copy of chunk k
    overlapped with the walk kernel of chunk k+1 (two streams).  Chunks are ranges of start vertices;
    a walker's random stream depends on (seed, start vertex, walk number) only, so the rows are
    the ones a single launch would give.  ``graph`` is a DeviceGraph or PartitionedGraph.
    out_host: int32 [>= W, walk_length+1], contiguous, pinned.
    Returns (walks_device [W, L+1] view, alive_device bool[W]); the call returns with the copy DONE
    (it synchronises the side stream), rows of dropped walkers are still in place (filter with alive)."""
    dev = graph.device
    start_t = _as_device_i32(start, dev)
    n_start, L1 = int(start_t.numel()), int(walk_length) + 1
    W, pitch = n_start * int(num_walks), (L1 + 7) // 8 * 8
    if (out_host.dtype != torch.int32 or out_host.dim() != 2 or out_host.shape[0] < W or out_host.shape[1] != L1
            or not out_host.is_contiguous() or out_host.device.type != "cpu"):
        raise ValueError(f"out must be a contiguous int32 host tensor of shape [>= {W}, {L1}]")
    if not out_host.is_pinned():
        raise ValueError("out must be pinned host memory (tensor.pin_memory()): the copy is asynchronous")
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    if chunk_walkers is None:
        # ~131 k walkers per chunk on small jobs (measured best on the 800 k-walker config); large jobs use 32
        # chunks so that every launch still fills the GPU (148 SMs x 6 CTAs x 256 lanes)
        chunk_walkers = max(1 << 17, W // 32)
    with torch.cuda.device(dev):
        # every buffer is allocated on the launching stream and outlives the side stream's work (the call
        # ends with side.synchronize()), so the caching allocator never sees a cross-stream tensor
        buf = torch.empty((W, pitch), dtype=torch.int32, device=dev)       # kernel output, pitch-padded rows
        rows = torch.empty((W, L1), dtype=torch.int32, device=dev)         # compact rows: D2H source, returned
        cur = torch.cuda.current_stream(dev)
        side = getattr(graph, "_copy_stream", None)
        if side is None:
            side = graph._copy_stream = torch.cuda.Stream(device=dev)
        per = max(1, int(chunk_walkers) // max(int(num_walks), 1))
        alive_parts = []
        for s0 in range(0, n_start, per):
            s1 = min(n_start, s0 + per)
            r0, r1 = s0 * int(num_walks), s1 * int(num_walks)
            _, a, _ = graph.walk(start_t[s0:s1], num_walks, walk_length, return_param, inout_param, seed, out=buf[r0:r1])
            alive_parts.append(a)
            # compaction stays on the launching stream: on the side stream it would queue behind the next
            # chunk's walk kernel for SM slots and serialise the pipeline (measured: 3.3 ms vs 2.9 ms)
            rows[r0:r1].copy_(buf[r0:r1, :L1])
            done = torch.cuda.Event()
            done.record(cur)
            side.wait_event(done)
            with torch.cuda.stream(side):
                out_host[r0:r1].copy_(rows[r0:r1], non_blocking=True)
        alive = torch.cat(alive_parts) if alive_parts else torch.zeros(0, dtype=torch.bool, device=dev)
        side.synchronize()
    return rows, alive


def walk_consts(return_param: float, inout_param: float, flags: int, has_ratio: bool = False) -> "_lib.WalkConsts":
    """Host-only: the sampling constants n2v_walk will use (no GPU needed)."""
    c = _lib.WalkConsts()
    _lib.check(_lib.load().n2v_walk_consts(float(return_param), float(inout_param), int(flags),
                                           1 if has_ratio else 0, C.byref(c)))
    return c


# ======================================================================================
# Vertex-partitioned CSR: every rank owns the arcs of a contiguous vertex range; peers read
# them over NVLink (CUDA-IPC-mapped pointers in n2v_graph_t.parts[]).
# ======================================================================================
class _RawBuffer(object):
    """Zero-copy torch view of a library-allocated (IPC-shareable) device buffer."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def _ipc_tensor(lib, shape, dtype: torch.dtype, device) -> Tuple[torch.Tensor, int]:
    typestr = {torch.int32: "<i4", torch.float64: "<f8"}[dtype]
    n = int(np.prod(shape)) if len(shape) else 1
    nbytes = max(n, 1) * (4 if dtype == torch.int32 else 8)
    ptr = C.c_void_p()
    _lib.check(lib.n2v_ipc_alloc(nbytes, C.byref(ptr)), "n2v_ipc_alloc")
    if n == 0:
        return torch.empty(shape, dtype=dtype, device=device), int(ptr.value)
    return torch.as_tensor(_RawBuffer(ptr.value, shape, typestr), device=device), int(ptr.value)


def _exchange_fds(my_fds, rank: int, world: int, group) -> Dict[int, list]:
    """Give every peer process its own copy of this rank's shareable-handle file descriptors
    (SCM_RIGHTS over Unix sockets) and collect theirs.  Returns {peer rank: [fds]}."""
    import socket
    import tempfile
    import threading
    import torch.distributed as dist
    path = os.path.join(tempfile.gettempdir(), f"n2v_ipc_{os.getpid()}_{rank}.sock")
    if os.path.exists(path):
        os.unlink(path)
    server = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    server.bind(path)
    server.listen(world)
    paths = [None] * world
    dist.all_gather_object(paths, path, group=group)

    def serve():
        for _ in range(world - 1):
            conn, _ = server.accept()
            socket.send_fds(conn, [b"n2v"], list(my_fds))
            conn.close()
    t = threading.Thread(target=serve, daemon=True)
    t.start()
    got = {}
    for p in range(world):
        if p == rank:
            continue
        c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        c.connect(paths[p])
        _, fds, _, _ = socket.recv_fds(c, 16, len(my_fds))
        c.close()
        got[p] = list(fds)
    t.join()
    server.close()
    os.unlink(path)
    return got


class PartitionedGraph(DeviceGraph):
    """One rank's part of a vertex-range-partitioned graph (BASELINE configs[4]).

    Rank r owns vertices [r*S, (r+1)*S), S = ceil(V / G): their arcs, alias records, hash sets
    and weights live in ITS HBM (IPC-shareable allocations); the 16-byte vertex headers are
    all-gathered (replicated, 16 B/vertex).  A walker started on rank r stays on rank r and
    dereferences remote parts with plain loads through NVLink/NVSwitch -- no walker migration,
    no collective in the walk.  Results are bit-identical to the replicated graph's.
    """

    @classmethod
    def from_local_arcs(cls, src, dst, weight, n_vertices: int, group=None, assume_symmetric: bool = False,
                        sum_mode: str = "naive", keep_weight: bool = True) -> "PartitionedGraph":
        """src/dst/weight: the arcs whose src lies in this rank's range (global ids).
        ``keep_weight=False`` frees the fp64 weights (8 B/arc) after the build when every part is
        unit-weight: the walk needs them only in its exact fallback, which then uses 1.0."""
        import torch.distributed as dist
        from . import dist as n2v_dist
        _lib.require_cuda()
        lib = _lib.load()
        rank, G = n2v_dist.world(group)
        if G > _lib.N2V_MAX_PARTS:
            raise ValueError(f"at most {_lib.N2V_MAX_PARTS} parts")
        device = torch.device("cuda", torch.cuda.current_device())
        S = (int(n_vertices) + G - 1) // G
        lo = rank * S
        g = cls()
        g.device, g.sum_mode, g.group, g.rank, g.n_parts, g.part_size = device, sum_mode, group, rank, G, S
        g.v_lo, g.v_hi = lo, min(int(n_vertices), lo + S)
        with torch.cuda.device(device):
            s = _as_device_i32(src, device) - lo
            d = _as_device_i32(dst, device)
            w = None if weight is None else _as_device_f64(weight, device)
            A = int(s.numel())
            stream = _lib.current_stream_ptr()
            g._ipc_ptrs = []
            vtx_local = torch.zeros((S, 4), dtype=torch.int32, device=device)
            g.col, p = _ipc_tensor(lib, (A,), torch.int32, device); g._ipc_ptrs.append(p)
            if keep_weight:
                g.weight, p = _ipc_tensor(lib, (A,), torch.float64, device); g._ipc_ptrs.append(p)
            else:
                g.weight = torch.empty(A, dtype=torch.float64, device=device); g._ipc_ptrs.append(0)
            nbytes = int(lib.n2v_csr_scratch_bytes(A, S))
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=device)
            flags = C.c_uint32(0)
            _lib.check(lib.n2v_csr_build(_lib.ptr(s), _lib.ptr(d), _lib.ptr(w), A, S, int(n_vertices),
                                         _lib.ptr(vtx_local), _lib.ptr(g.col), _lib.ptr(g.weight), None,
                                         _lib.ptr(scratch), nbytes, C.byref(flags), stream), "n2v_csr_build")
            del scratch, s, d, w
            n_buckets = int(lib.n2v_hash_buckets_bound(A, S))
            g.hash, p = _ipc_tensor(lib, (n_buckets, 8), torch.int32, device); g._ipc_ptrs.append(p)
            _lib.check(lib.n2v_hash_build(_lib.ptr(vtx_local), _lib.ptr(g.col), S, A, _lib.ptr(g.hash), n_buckets,
                                          stream), "n2v_hash_build")
            # replicate the headers (16 B/vertex) so landing-vertex fields and walk starts are local
            g.vtx = torch.empty((G * S, 4), dtype=torch.int32, device=device)
            if G > 1:
                dist.all_gather_into_tensor(g.vtx, vtx_local, group=group)
            else:
                g.vtx.copy_(vtx_local)
            g.arcs, p = _ipc_tensor(lib, (A, 8), torch.int32, device); g._ipc_ptrs.append(p)
            probs = torch.empty(A, dtype=torch.float64, device=device)
            work = torch.empty(A, dtype=torch.int32, device=device)
            n_zero = C.c_int64(0)
            _lib.check(lib.n2v_alias_build(_lib.ptr(vtx_local), _lib.ptr(g.vtx), _lib.ptr(g.col), _lib.ptr(g.weight), S,
                                           A, _lib.SUM_MODE[sum_mode], None, _lib.ptr(probs), _lib.ptr(g.arcs),
                                           _lib.ptr(work), C.byref(n_zero), stream), "n2v_alias_build")
            g.vtx[lo:lo + S] = vtx_local          # own wsum
            del probs, work
            torch.cuda.synchronize()
            # graph-wide flags: AND over parts; symmetry cannot be checked part-locally
            bits = (_lib.GRAPH_UNIT_WEIGHT, _lib.GRAPH_SYMMETRIC, _lib.GRAPH_SIMPLE)
            fl = torch.tensor([1 if int(flags.value) & b else 0 for b in bits], dtype=torch.int32, device=device)
            totals = torch.tensor([A], dtype=torch.int64, device=device)
            if G > 1:
                dist.all_reduce(fl, op=dist.ReduceOp.MIN, group=group)      # AND over parts
                dist.all_reduce(totals, op=dist.ReduceOp.SUM, group=group)
            g.flags = sum(b for b, on in zip(bits, fl.tolist()) if on) & ~_lib.GRAPH_SYMMETRIC
            if assume_symmetric and (g.flags & _lib.GRAPH_SIMPLE):
                g.flags |= _lib.GRAPH_SYMMETRIC
            g.n_vertices, g.n_arcs, g.n_local_arcs = int(n_vertices), int(totals.item()), A
            if not keep_weight:
                if g.flags & _lib.GRAPH_UNIT_WEIGHT:
                    g.weight = None                      # every part is unit-weight: nobody reads them
                else:                                    # weights matter after all: move them to a shareable buffer
                    shared_w, p = _ipc_tensor(lib, (A,), torch.float64, device)
                    shared_w.copy_(g.weight)
                    g.weight, g._ipc_ptrs[1] = shared_w, p
                    torch.cuda.synchronize()
            # export shareable handles, hand the peers their own copies of the fds, map their parts
            import struct
            handles, my_fds = [], []
            own = {"arcs": g._ipc_ptrs[3], "col": g._ipc_ptrs[0], "weight": g._ipc_ptrs[1], "hash": g._ipc_ptrs[2]}
            shared = ("arcs", "col", "hash") if g.weight is None else ("arcs", "col", "weight", "hash")
            for name in shared:
                buf = C.create_string_buffer(64)
                _lib.check(lib.n2v_ipc_export(C.c_void_p(own[name]), buf), "n2v_ipc_export")
                handles.append(bytes(buf.raw))
                my_fds.append(struct.unpack_from("<i", buf.raw, 0)[0])
            all_handles = [None] * G
            peer_fds = {}
            if G > 1:
                dist.all_gather_object(all_handles, handles, group=group)
                peer_fds = _exchange_fds(my_fds, rank, G, group)
            else:
                all_handles[0] = handles
            st = _lib.Graph()
            st.n_vertices, st.n_arcs, st.flags = g.n_vertices, g.n_arcs, g.flags
            st.n_parts, st.part_size = G, S
            g._peer_ptrs = []
            for p_idx in range(G):
                st.parts[p_idx].vtx = g.vtx.data_ptr() + p_idx * S * 16
                if p_idx == rank:
                    ptrs = {"arcs": g.arcs.data_ptr(), "col": g.col.data_ptr(), "hash": g.hash.data_ptr(),
                            "weight": None if g.weight is None else g.weight.data_ptr()}
                else:
                    ptrs = {"weight": None}
                    for name, h, fd in zip(shared, all_handles[p_idx], peer_fds[p_idx]):
                        local_h = struct.pack("<i", fd) + h[4:]          # the fd as it is known in THIS process
                        out = C.c_void_p()
                        _lib.check(lib.n2v_ipc_open(local_h, C.byref(out)), "n2v_ipc_open")
                        os.close(fd)
                        ptrs[name] = int(out.value)
                        g._peer_ptrs.append(int(out.value))
                st.parts[p_idx].arcs, st.parts[p_idx].col = ptrs["arcs"], ptrs["col"]
                st.parts[p_idx].weight, st.parts[p_idx].hash = ptrs["weight"], ptrs["hash"]
            g._struct = st
            for fd in my_fds:
                os.close(fd)
            if G > 1:
                dist.barrier(group=group)
        return g

    def localize(self, reserve_bytes: int = 0) -> bool:
        """Turn the vertex-partitioned graph into a REPLICATED one assembled from the partitioned
        build, when it fits: every rank copies its peers' arc records and hash sets (40 B per arc;
        the peers' parts are already mapped here, so the copies are plain device-to-device
        transfers over NVLink) and re-points ``parts[]`` to the copies.  The walk kernel is
        unchanged -- same part-relative records, same results bit for bit -- but no gather
        crosses NVLink any more ("a replicated CSR where it fits in 180 GB", BASELINE north_star).
        ``col`` / ``weight`` (exact fallback only) stay remote.  Collective: every rank calls it;
        returns False and changes nothing unless EVERY rank has room for the copies plus
        ``reserve_bytes``."""
        import torch.distributed as dist
        if self.n_parts == 1 or getattr(self, "_local_parts", None):
            return True
        dev, G, rank = self.device, self.n_parts, self.rank
        with torch.cuda.device(dev):
            sizes = torch.zeros((G, 2), dtype=torch.int64, device=dev)
            sizes[rank, 0], sizes[rank, 1] = int(self.arcs.shape[0]), int(self.hash.shape[0])
            dist.all_reduce(sizes, group=self.group)
            sizes = sizes.cpu().tolist()
            need = sum((a + b) * 32 for p, (a, b) in enumerate(sizes) if p != rank)
            torch.cuda.empty_cache()
            free, _ = torch.cuda.mem_get_info(dev)
            ok = torch.tensor([1 if free - need - int(reserve_bytes) > (2 << 30) else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                return False
            local = {}
            for p, (a, b) in enumerate(sizes):
                if p == rank:
                    continue
                arcs = torch.empty((a, 8), dtype=torch.int32, device=dev)
                hsh = torch.empty((b, 8), dtype=torch.int32, device=dev)
                if a:
                    arcs.copy_(torch.as_tensor(_RawBuffer(self._struct.parts[p].arcs, (a, 8), "<i4"), device=dev))
                if b:
                    hsh.copy_(torch.as_tensor(_RawBuffer(self._struct.parts[p].hash, (b, 8), "<i4"), device=dev))
                local[p] = (arcs, hsh)
            torch.cuda.synchronize()
            dist.barrier(group=self.group)             # nobody frees a part a peer is still copying
            for p, (arcs, hsh) in local.items():
                self._struct.parts[p].arcs = arcs.data_ptr()
                self._struct.parts[p].hash = hsh.data_ptr()
            self._local_parts = local
        return True

    def start_vertices(self) -> torch.Tensor:
        """This rank's start vertices (global ids): its own range, out-degree > 0."""
        own = self.vtx[self.v_lo:self.v_hi, 1]
        return (torch.nonzero(own != 0).view(-1) + self.v_lo).to(torch.int32)

    def nbytes(self) -> int:
        return sum(int(t.numel()) * t.element_size() for t in (self.vtx, self.arcs, self.col, self.weight, self.hash)
                   if t is not None)

    def close(self) -> None:
        """Unmap peers' parts and free this rank's shareable buffers (after a barrier)."""
        import torch.distributed as dist
        lib = _lib.load()
        torch.cuda.synchronize()
        for p in getattr(self, "_peer_ptrs", []):
            lib.n2v_ipc_close(C.c_void_p(p))
        self._peer_ptrs = []
        if self.n_parts > 1 and dist.is_initialized():
            dist.barrier(group=self.group)
        self.arcs = self.col = self.weight = self.hash = None
        self._local_parts = None
        for p in getattr(self, "_ipc_ptrs", []):
            if p:
                lib.n2v_ipc_free(C.c_void_p(p))
        self._ipc_ptrs = []
