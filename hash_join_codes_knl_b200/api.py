"""Host-side mirror of the reference's join interface over the C ABI (include/hjb200.h).

Reference shape: main() loads four uint32 columns, fills one info_t_hj per thread (hj.h:1-72)
and runs run()/run_hj() (npj.cpp:769, phj.cpp:1646, cpra2.cpp:1697); the result is the
(join_keys, join_outer_vals, join_inner_vals) columns and join_tuples.  Here an Engine is one
GPU; `inner` = R = build side, `outer` = S = probe side, as in the reference."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Gen, Opts, Recv, Rel, Result, Split


class HjbError(RuntimeError):
    pass


class HjbCapacityError(HjbError):
    """hjb_cpra_finish: an owner's receive buffer was too small (HJB_E_CAPACITY); `largest` = (R rows, S rows)
    the fullest owner received."""

    def __init__(self, msg, largest):
        super().__init__(msg)
        self.largest = largest


class _CudaArray:
    """Zero-copy view of library-owned device memory for torch.as_tensor (int32 bit pattern)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class JoinResult:
    """count + the three uint64 checksums + the dense result columns."""

    def __init__(self, res, engine):
        self.count, self.sum_key = int(res.count), int(res.sum_key)
        self.sum_outer, self.sum_inner = int(res.sum_outer), int(res.sum_inner)
        self.seconds, self.seconds_e2e = float(res.seconds), float(res.seconds_e2e)
        self.phase_ms = [float(x) for x in res.phase_ms]
        self.kernel_launches, self.partitions = int(res.kernel_launches), int(res.partitions)
        self.rows_on_device = bool(res.rows_on_device)
        self._ptrs = (res.keys, res.outer_vals, res.inner_vals)
        self._engine = engine

    def checks(self):
        return (self.count, self.sum_key, self.sum_outer, self.sum_inner)

    @property
    def materialized(self):
        return self._ptrs[0] is not None or self.count == 0

    def rows_numpy(self):
        """(keys, outer_vals, inner_vals) as host uint32 arrays (copies)."""
        n = self.count
        if n == 0:
            return tuple(np.empty(0, np.uint32) for _ in range(3))
        if self._ptrs[0] is None:
            raise HjbError("join ran with materialize=0")
        if self.rows_on_device:
            import torch
            return tuple(torch.as_tensor(_CudaArray(p, n), device=f"cuda:{self._engine.device}").cpu().numpy()
                         .view(np.uint32).copy() for p in self._ptrs)
        return tuple(np.ctypeslib.as_array((C.c_uint32 * n).from_address(p)).copy() for p in self._ptrs)

    def rows_torch(self):
        """(keys, outer_vals, inner_vals) as int32 CUDA tensors aliasing the engine's buffers
        (valid until the next join on the engine)."""
        import torch
        if not self.rows_on_device:
            raise HjbError("rows are in host memory")
        n = self.count
        dev = f"cuda:{self._engine.device}"
        if n == 0:
            return tuple(torch.empty(0, dtype=torch.int32, device=dev) for _ in range(3))
        return tuple(torch.as_tensor(_CudaArray(p, n), device=dev) for p in self._ptrs)


def _is_torch(x):
    return type(x).__module__.startswith("torch")


_own_stream_engines = 0        # Engines that run on a stream of their own (not torch's current stream)


def _col(x):
    """-> (pointer, tuples, on_device, keepalive)"""
    if _is_torch(x):
        assert x.dim() == 1 and x.is_contiguous() and x.element_size() == 4, "columns are contiguous 1-D 32-bit"
        if x.is_cuda and _own_stream_engines:
            # the tensor may still be being written by kernels queued on torch's stream, which the engine's
            # non-blocking stream does not wait for
            import torch
            torch.cuda.current_stream(x.device).synchronize()
        return x.data_ptr(), x.numel(), x.is_cuda, x
    a = np.ascontiguousarray(x)
    assert a.ndim == 1 and a.dtype.itemsize == 4, "columns are 1-D 32-bit"
    return a.ctypes.data, a.size, False, a


class Engine:
    """One GPU: wraps hjb_ctx.  Raises if libhjb200.so or a CUDA device is missing."""

    def __init__(self, device=0, use_torch_stream=False):
        self._lib = _lib.load()
        self.device = int(device)
        self._ctx = C.c_void_p()
        rc = self._lib.hjb_create(self.device, C.byref(self._ctx))
        if rc != 0:
            msg = self._lib.hjb_last_error(None).decode()
            self._ctx = None
            raise HjbError(f"hjb_create({device}) failed ({rc}): {msg}")
        global _own_stream_engines
        self._own_stream = not use_torch_stream
        if use_torch_stream:
            import torch
            self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        else:
            _own_stream_engines += 1

    def close(self):
        global _own_stream_engines
        if getattr(self, "_ctx", None):
            self._lib.hjb_destroy(self._ctx)
            self._ctx = None
            if getattr(self, "_own_stream", False):
                _own_stream_engines -= 1

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise HjbError(f"{what} failed ({rc}): {self._lib.hjb_last_error(self._ctx).decode()}")

    def set_stream(self, cuda_stream):
        self._check(self._lib.hjb_set_stream(self._ctx, C.c_void_p(int(cuda_stream))), "hjb_set_stream")

    def synchronize(self):
        self._check(self._lib.hjb_synchronize(self._ctx), "hjb_synchronize")

    def set_profiling(self, on=True):
        self._check(self._lib.hjb_set_profiling(self._ctx, int(on)), "hjb_set_profiling")

    def kernel_times(self):
        """{kernel name: (total ms, launches)} of the last whole join (needs set_profiling(True))."""
        ms = (C.c_float * 16)()
        n = (C.c_uint32 * 16)()
        kinds = self._lib.hjb_kernel_times(self._ctx, ms, n, 16)
        if kinds < 0:
            self._check(kinds, "hjb_kernel_times")
        return {self._lib.hjb_kernel_name(k).decode(): (float(ms[k]), int(n[k])) for k in range(kinds)}

    @staticmethod
    def _opts(materialize=True, seed=0, npj_load=0.0, radix_bits=(), part_tuples=0, out_capacity=0):
        o = Opts()
        o.materialize, o.seed, o.npj_load = int(bool(materialize)), int(seed), float(npj_load)
        for i, b in enumerate(radix_bits):
            o.radix_bits[i] = int(b)
        o.part_tuples, o.out_capacity = int(part_tuples), int(out_capacity)
        return o

    def _rels(self, inner, outer):
        keep = []
        rels = []
        dev = None
        for keys, vals in (inner, outer):
            kp, kn, kd, ka = _col(keys)
            vp, vn, vd, va = _col(vals)
            if kn != vn:
                raise HjbError("key and payload columns differ in length")
            if kd != vd or (dev is not None and dev != kd):
                raise HjbError("all four columns must live on the same side (host or device)")
            dev = kd
            keep += [ka, va]
            rels.append(Rel(kp if kn else None, vp if kn else None, kn))
        return rels[0], rels[1], dev, keep

    def _join(self, algo, inner, outer, opts):
        R, S, on_dev, keep = self._rels(inner, outer)
        fn = getattr(self._lib, f"hjb_{algo}_{'device' if on_dev else 'host'}")
        res = Result()
        o = self._opts(**opts)
        self._check(fn(self._ctx, C.byref(R), C.byref(S), C.byref(o), C.byref(res)), fn.__name__)
        del keep
        return JoinResult(res, self)

    def npj(self, inner, outer, **opts):
        """Non-partitioned join (npj.cpp).  inner / outer = (keys, vals), both numpy (host entry
        point, copies in and out) or both CUDA tensors (device entry point)."""
        return self._join("npj", inner, outer, opts)

    def phj(self, inner, outer, **opts):
        """Radix-partitioned join (phj.cpp)."""
        return self._join("phj", inner, outer, opts)

    # ---- CPRA pieces (cpra2.cpp:1697-1986); hash_join_codes_knl_b200.cpra strings them together
    def cpra_split(self, inner_chunk, outer_chunk, ngpus, **opts):
        R, S, on_dev, keep = self._rels(inner_chunk, outer_chunk)
        if not on_dev:
            raise HjbError("cpra_split takes device columns")
        sp = Split()
        o = self._opts(**opts)
        self._check(self._lib.hjb_cpra_split(self._ctx, C.byref(R), C.byref(S), int(ngpus), C.byref(o), C.byref(sp)),
                    "hjb_cpra_split")
        import torch
        dev = f"cuda:{self.device}"

        def view(p, n):
            return torch.as_tensor(_CudaArray(p, n), device=dev) if n else torch.empty(0, dtype=torch.int32, device=dev)
        return {"r_keys": view(sp.r_keys, R.tuples), "r_vals": view(sp.r_vals, R.tuples),
                "s_keys": view(sp.s_keys, S.tuples), "s_vals": view(sp.s_vals, S.tuples),
                "r_offsets": [int(sp.r_offsets[g]) for g in range(ngpus + 1)],
                "s_offsets": [int(sp.s_offsets[g]) for g in range(ngpus + 1)], "ms": float(sp.ms), "_keep": keep}

    def cpra_join_local(self, inner_recv, outer_recv, gpu, ngpus, **opts):
        R, S, on_dev, keep = self._rels(inner_recv, outer_recv)
        if not on_dev:
            raise HjbError("cpra_join_local takes device columns")
        res = Result()
        o = self._opts(**opts)
        self._check(self._lib.hjb_cpra_join_local(self._ctx, C.byref(R), C.byref(S), int(gpu), int(ngpus), C.byref(o),
                                                  C.byref(res)), "hjb_cpra_join_local")
        return JoinResult(res, self)

    # ---- CPRA with the exchange fused into the GPU-assign pass (peer stores over NVLink)
    def cpra_recv_alloc(self, r_capacity, s_capacity):
        """(Re)allocates this GPU's receive buffers; returns {"ptrs": 4 device pointers, "ipc": 4 x 64-byte handles}."""
        rv = Recv()
        self._check(self._lib.hjb_cpra_recv_alloc(self._ctx, int(r_capacity), int(s_capacity), C.byref(rv)),
                    "hjb_cpra_recv_alloc")
        return {"ptrs": [rv.r_keys, rv.r_vals, rv.s_keys, rv.s_vals], "ipc": [bytes(rv.ipc[i]) for i in range(4)],
                "r_capacity": int(rv.r_capacity), "s_capacity": int(rv.s_capacity)}

    def ipc_open(self, handle):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        self._check(self._lib.hjb_ipc_open(self._ctx, buf, C.byref(p)), "hjb_ipc_open")
        return p.value

    def ipc_close(self, ptr):
        self._check(self._lib.hjb_ipc_close(self._ctx, C.c_void_p(ptr)), "hjb_ipc_close")

    def cpra_count(self, inner_chunk, outer_chunk, ngpus, **opts):
        R, S, on_dev, keep = self._rels(inner_chunk, outer_chunk)
        if not on_dev:
            raise HjbError("cpra_count takes device columns")
        rc, sc = (C.c_uint64 * 64)(), (C.c_uint64 * 64)()
        o = self._opts(**opts)
        self._check(self._lib.hjb_cpra_count(self._ctx, C.byref(R), C.byref(S), int(ngpus), C.byref(o), rc, sc),
                    "hjb_cpra_count")
        self._pending_keep = keep          # the chunks must outlive cpra_scatter_peer
        return [int(rc[g]) for g in range(ngpus)], [int(sc[g]) for g in range(ngpus)]

    def cpra_scatter_peer(self, ngpus, peer_ptrs, r_base, s_base):
        """peer_ptrs[c][g]: column c (r_keys, r_vals, s_keys, s_vals) of owner g as mapped into this process."""
        cols = [(C.c_void_p * ngpus)(*[C.c_void_p(p) for p in peer_ptrs[c]]) for c in range(4)]
        rb = (C.c_uint64 * ngpus)(*[int(x) for x in r_base])
        sb = (C.c_uint64 * ngpus)(*[int(x) for x in s_base])
        ms = C.c_float()
        self._check(self._lib.hjb_cpra_scatter_peer(self._ctx, int(ngpus), cols[0], cols[1], cols[2], cols[3], rb, sb,
                                                    C.byref(ms)), "hjb_cpra_scatter_peer")
        self._pending_keep = None
        return float(ms.value)

    # ---- the stream-ordered CPRA step (include/hjb200.h: hjb_cpra_bind ... hjb_cpra_finish)
    def cpra_bind(self, gpu, ngpus, peer_ptrs, r_capacity, s_capacity):
        cols = [(C.c_void_p * ngpus)(*[C.c_void_p(p) for p in peer_ptrs[c]]) for c in range(4)]
        self._check(self._lib.hjb_cpra_bind(self._ctx, int(gpu), int(ngpus), cols[0], cols[1], cols[2], cols[3],
                                            int(r_capacity), int(s_capacity)), "hjb_cpra_bind")

    def cpra_count_async(self, inner_chunk, outer_chunk, counts_dev, **opts):
        """counts_dev: int64 CUDA tensor of 2*ngpus elements (the all-gather's input).  Host columns (numpy /
        pinned torch CPU tensors) go through hjb_cpra_count_async_host, which copies them in on the stream."""
        R, S, on_dev, keep = self._rels(inner_chunk, outer_chunk)
        o = self._opts(**opts)
        fn = self._lib.hjb_cpra_count_async if on_dev else self._lib.hjb_cpra_count_async_host
        self._check(fn(self._ctx, C.byref(R), C.byref(S), C.byref(o), C.c_void_p(counts_dev.data_ptr())),
                    "hjb_cpra_count_async" + ("" if on_dev else "_host"))
        self._pending_keep = keep          # the chunks must outlive the scatter
        self._step_from_host = not on_dev

    def cpra_scatter_async(self, matrix_dev):
        self._check(self._lib.hjb_cpra_scatter_async(self._ctx, C.c_void_p(matrix_dev.data_ptr())), "hjb_cpra_scatter_async")

    def cpra_join_async(self, r_expect=0, s_expect=0, **opts):
        o = self._opts(**opts)
        self._check(self._lib.hjb_cpra_join_async(self._ctx, C.byref(o), int(r_expect), int(s_expect)), "hjb_cpra_join_async")

    def cpra_sums_dev(self):
        """int64 CUDA tensor aliasing this GPU's (count, sum_key, sum_outer, sum_inner) of the enqueued join."""
        return self.device_view64(self._lib.hjb_cpra_sums_dev(self._ctx), 4)

    def cpra_finish(self):
        """-> (JoinResult, (R rows, S rows) received here, (R rows, S rows) received by the fullest owner)"""
        res = Result()
        got, big = (C.c_uint64 * 2)(), (C.c_uint64 * 2)()
        fn = self._lib.hjb_cpra_finish_host if getattr(self, "_step_from_host", False) else self._lib.hjb_cpra_finish
        rc = fn(self._ctx, C.byref(res), got, big)
        self._pending_keep = None
        if rc == _lib.HJB_E_CAPACITY:
            raise HjbCapacityError(self._lib.hjb_last_error(self._ctx).decode(), (int(big[0]), int(big[1])))
        self._check(rc, "hjb_cpra_finish")
        return JoinResult(res, self), (int(got[0]), int(got[1])), (int(big[0]), int(big[1]))

    # ---- the staged exchange (include/hjb200.h: hjb_cpra_stage_*)
    def cpra_stage_plan(self, ngpus, r_expect, s_expect, **opts):
        """-> (abits, bbits, big_fill) or None when two passes do not suffice (use the fused path)"""
        a, b, f = C.c_int(), C.c_int(), C.c_int()
        o = self._opts(**opts)
        rc = self._lib.hjb_cpra_stage_plan(self._ctx, int(ngpus), int(r_expect), int(s_expect), C.byref(o), C.byref(a), C.byref(b),
                                           C.byref(f))
        return None if rc != 0 else (int(a.value), int(b.value), int(f.value))

    def cpra_stage_count_async(self, inner_chunk, outer_chunk, abits, counts_dev, nparts=1, **opts):
        """counts_dev: int64 CUDA tensor of 2 * 2^abits elements (the all-gather's input); nparts: the runs leave, and are
        processed by their owners, in this many parts (ranges of sub-partitions)"""
        R, S, on_dev, keep = self._rels(inner_chunk, outer_chunk)
        if not on_dev:
            raise HjbError("cpra_stage_count_async takes device columns")
        o = self._opts(**opts)
        self._check(self._lib.hjb_cpra_stage_count_async(self._ctx, C.byref(R), C.byref(S), C.byref(o), int(abits), int(nparts),
                                                         C.c_void_p(counts_dev.data_ptr())), "hjb_cpra_stage_count_async")
        self._pending_keep = keep
        self._step_from_host = False

    def cpra_stage_scatter_async(self, matrix_dev, rel):
        self._check(self._lib.hjb_cpra_stage_scatter_async(self._ctx, C.c_void_p(matrix_dev.data_ptr()), int(rel)),
                    "hjb_cpra_stage_scatter_async")

    def cpra_stage_copy_async(self, rel, cuda_stream=None, part=0):
        """cuda_stream: raw cudaStream_t of a side stream that already waits for the scatter (None: the engine's stream)"""
        self._check(self._lib.hjb_cpra_stage_copy_async(self._ctx, int(rel), int(part), C.c_void_p(int(cuda_stream) if cuda_stream else None)),
                    "hjb_cpra_stage_copy_async")

    def cpra_stage_local_async(self, bbits, big_fill, rel, part=0, **opts):
        o = self._opts(**opts)
        self._check(self._lib.hjb_cpra_stage_local_async(self._ctx, C.byref(o), int(bbits), int(big_fill), int(rel), int(part)),
                    "hjb_cpra_stage_local_async")

    # ---- heavy-hitter handling (include/hjb200.h: hjb_cpra_split_hot / _select_hot / _hot_join)
    def cpra_split_hot(self, outer_chunk, hot_keys_dev):
        """-> ((cold keys, cold vals), (hot keys, hot vals)): int32 CUDA tensors aliasing context memory"""
        kp, kn, kd, ka = _col(outer_chunk[0])
        vp, vn, vd, va = _col(outer_chunk[1])
        assert kd and vd and kn == vn
        S = Rel(kp if kn else None, vp if kn else None, kn)
        cold, hot = Rel(), Rel()
        self._check(self._lib.hjb_cpra_split_hot(self._ctx, C.byref(S), C.c_void_p(hot_keys_dev.data_ptr()), int(hot_keys_dev.numel()),
                                                 C.byref(cold), C.byref(hot)), "hjb_cpra_split_hot")
        view = lambda r: (self.device_view(r.keys, r.tuples), self.device_view(r.vals, r.tuples))
        return view(cold), view(hot)

    def cpra_select_hot(self, inner_chunk, hot_keys_dev, keys_out, vals_out):
        """this chunk's build tuples with hot keys -> keys_out / vals_out (int32 CUDA tensors); returns how many were found"""
        kp, kn, kd, ka = _col(inner_chunk[0])
        vp, vn, vd, va = _col(inner_chunk[1])
        assert kd and vd and kn == vn
        R = Rel(kp if kn else None, vp if kn else None, kn)
        found = C.c_uint64()
        self._check(self._lib.hjb_cpra_select_hot(self._ctx, C.byref(R), C.c_void_p(hot_keys_dev.data_ptr()), int(hot_keys_dev.numel()),
                                                  C.c_void_p(keys_out.data_ptr()), C.c_void_p(vals_out.data_ptr()),
                                                  int(keys_out.numel()), C.byref(found)), "hjb_cpra_select_hot")
        return int(found.value)

    def cpra_hot_join(self, hot_outer, hot_inner):
        S = Rel(hot_outer[0].data_ptr() if hot_outer[0].numel() else None, hot_outer[1].data_ptr() if hot_outer[0].numel() else None,
                hot_outer[0].numel())
        R = Rel(hot_inner[0].data_ptr() if hot_inner[0].numel() else None, hot_inner[1].data_ptr() if hot_inner[0].numel() else None,
                hot_inner[0].numel())
        self._check(self._lib.hjb_cpra_hot_join(self._ctx, C.byref(S), C.byref(R)), "hjb_cpra_hot_join")
        self._hot_keep = (hot_outer, hot_inner)

    def host_register(self, array):
        """page-locks a caller-owned numpy array in place (hjb_host_register); undo with host_unregister"""
        self._check(self._lib.hjb_host_register(C.c_void_p(array.ctypes.data), array.nbytes), "hjb_host_register")

    def host_unregister(self, array):
        self._check(self._lib.hjb_host_unregister(C.c_void_p(array.ctypes.data)), "hjb_host_unregister")

    def device_view64(self, ptr, n):
        import torch
        arr = type("A", (), {"__cuda_array_interface__": {"shape": (int(n),), "typestr": "<i8", "data": (int(ptr), False),
                                                          "version": 2, "strides": None}})()
        return torch.as_tensor(arr, device=f"cuda:{self.device}")

    def device_view(self, ptr, n):
        """int32 CUDA tensor aliasing n elements of library-owned device memory."""
        import torch
        dev = f"cuda:{self.device}"
        return torch.as_tensor(_CudaArray(ptr, n), device=dev) if n else torch.empty(0, dtype=torch.int32, device=dev)

    # ---- single kernels, for comparing intermediate products with the oracle
    def hash_factor(self, seed, which):
        return int(self._lib.hjb_hash_factor(int(seed), int(which)))

    def histogram(self, keys_dev, factor, shift, bits):
        kp, n, on_dev, _ = _col(keys_dev)
        assert on_dev
        counts = np.zeros(1 << bits, np.uint32)
        self._check(self._lib.hjb_histogram(self._ctx, kp, n, counts.ctypes.data_as(_lib.u32p), int(factor), shift, bits),
                    "hjb_histogram")
        return counts

    def partition_pass(self, keys_dev, vals_dev, factor, shift, bits, parent_offsets=None):
        import torch
        kp, n, on_dev, _ = _col(keys_dev)
        vp, _, _, _ = _col(vals_dev)
        assert on_dev
        ko, vo = torch.empty_like(keys_dev), torch.empty_like(vals_dev)
        child = np.zeros((1 << (shift + bits)) + 1, np.uint32)
        par = None
        if parent_offsets is not None:
            par = np.ascontiguousarray(parent_offsets, dtype=np.uint32)
        self._check(self._lib.hjb_partition_pass(self._ctx, kp, vp, n,
                                                 par.ctypes.data_as(_lib.u32p) if par is not None else None,
                                                 ko.data_ptr(), vo.data_ptr(), child.ctypes.data_as(_lib.u32p),
                                                 int(factor), shift, bits), "hjb_partition_pass")
        return ko, vo, child

    def npj_build(self, keys_dev, vals_dev, buckets, factor):
        import torch
        kp, n, on_dev, _ = _col(keys_dev)
        vp, _, _, _ = _col(vals_dev)
        assert on_dev
        table = torch.empty(buckets * 4, dtype=torch.int64, device=keys_dev.device)
        self._check(self._lib.hjb_npj_build(self._ctx, kp, vp, n, table.data_ptr(), int(buckets), int(factor)),
                    "hjb_npj_build")
        return table

    def generate(self, kind, tuples, domain, seed, order_seed, payload_factor, first=0, total=None, theta=1.0,
                 selectivity=1.0):
        """Device generator (csrc/gen.cu).  Returns (keys, vals) int32 CUDA tensors."""
        import torch
        dev = f"cuda:{self.device}"
        keys = torch.empty(tuples, dtype=torch.int32, device=dev)
        vals = torch.empty(tuples, dtype=torch.int32, device=dev)
        g = Gen(int(kind), int(tuples), int(domain), int(first), int(total if total is not None else tuples),
                int(seed), int(order_seed), int(payload_factor), 0, float(theta), float(selectivity))
        self._check(self._lib.hjb_generate(self._ctx, C.byref(g), keys.data_ptr(), vals.data_ptr()), "hjb_generate")
        self.synchronize()
        return keys, vals

    def rows_fingerprint(self, keys_dev, outer_vals_dev, inner_vals_dev):
        """order-independent (sum, xor) fingerprint of result rows held in device columns"""
        cols = [_col(c) for c in (keys_dev, outer_vals_dev, inner_vals_dev)]
        assert all(c[2] for c in cols) and len({c[1] for c in cols}) == 1
        fp = (C.c_uint64 * 2)()
        self._check(self._lib.hjb_rows_fingerprint(self._ctx, cols[0][0], cols[1][0], cols[2][0], cols[0][1], fp),
                    "hjb_rows_fingerprint")
        return int(fp[0]), int(fp[1])

    def column_sum(self, col_dev):
        p, n, on_dev, _ = _col(col_dev)
        assert on_dev
        s = C.c_uint64()
        self._check(self._lib.hjb_column_sum(self._ctx, p, n, C.byref(s)), "hjb_column_sum")
        return int(s.value)


def rows_fingerprint_numpy(keys, outer_vals, inner_vals):
    """numpy mirror of hjb_rows_fingerprint (csrc/gen.cu k_rows_fingerprint): (sum, xor) of
    splitmix64((key | outer << 32) ^ splitmix64(inner)) over the rows, uint64 wrap-around."""
    def mix(z):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))
    with np.errstate(over="ignore"):
        k, o, i = (np.ascontiguousarray(c).view(np.uint32).astype(np.uint64) for c in (keys, outer_vals, inner_vals))
        z = mix((k | (o << np.uint64(32))) ^ mix(i))
        return int(z.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(z)) if z.size else 0


def relation_write(directory, outer, keys, vals):
    """write.cpp:1824-1865 format: <dir>/{i|o}k_<n>.txt and {i|o}v_<n>.txt, raw uint32."""
    k = np.ascontiguousarray(keys).view(np.uint32)
    v = np.ascontiguousarray(vals).view(np.uint32)
    rc = _lib.load().hjb_relation_write(str(directory).encode(), int(bool(outer)), k.size,
                                        k.ctypes.data_as(_lib.u32p), v.ctypes.data_as(_lib.u32p))
    if rc != 0:
        raise HjbError(f"hjb_relation_write failed ({rc})")


def relation_read(directory, outer, tuples):
    k, v = np.empty(tuples, np.uint32), np.empty(tuples, np.uint32)
    rc = _lib.load().hjb_relation_read(str(directory).encode(), int(bool(outer)), tuples,
                                       k.ctypes.data_as(_lib.u32p), v.ctypes.data_as(_lib.u32p))
    if rc != 0:
        raise HjbError(f"hjb_relation_read failed ({rc}): missing file or size != 4*tuples")
    return k, v
