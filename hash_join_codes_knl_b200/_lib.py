"""Loads libhjb200.so and declares the C ABI of include/hjb200.h for ctypes.  There is no
fallback: if the library is missing or fails to load this raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhjb200.so")

HJB_E_CAPACITY = -6

u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


class Rel(C.Structure):
    _fields_ = [("keys", C.c_void_p), ("vals", C.c_void_p), ("tuples", C.c_uint64)]


class Opts(C.Structure):
    _fields_ = [("materialize", C.c_int), ("seed", C.c_uint32), ("npj_load", C.c_double),
                ("radix_bits", C.c_int * 4), ("part_tuples", C.c_uint32), ("out_capacity", C.c_uint64),
                ("reserved", C.c_int * 8)]


class Result(C.Structure):
    _fields_ = [("count", C.c_uint64), ("sum_key", C.c_uint64), ("sum_outer", C.c_uint64),
                ("sum_inner", C.c_uint64), ("keys", C.c_void_p), ("outer_vals", C.c_void_p),
                ("inner_vals", C.c_void_p), ("rows_on_device", C.c_int), ("seconds", C.c_double),
                ("seconds_e2e", C.c_double), ("phase_ms", C.c_float * 8), ("kernel_launches", C.c_uint32),
                ("partitions", C.c_uint32)]


class Split(C.Structure):
    _fields_ = [("r_keys", C.c_void_p), ("r_vals", C.c_void_p), ("s_keys", C.c_void_p), ("s_vals", C.c_void_p),
                ("r_offsets", C.c_uint64 * 65), ("s_offsets", C.c_uint64 * 65), ("ms", C.c_float)]


class Recv(C.Structure):
    _fields_ = [("r_keys", C.c_void_p), ("r_vals", C.c_void_p), ("s_keys", C.c_void_p), ("s_vals", C.c_void_p),
                ("r_capacity", C.c_uint64), ("s_capacity", C.c_uint64), ("ipc", (C.c_ubyte * 64) * 4)]


class Gen(C.Structure):
    _fields_ = [("kind", C.c_int), ("tuples", C.c_uint64), ("domain", C.c_uint64), ("first", C.c_uint64),
                ("total", C.c_uint64), ("seed", C.c_uint32), ("order_seed", C.c_uint32), ("payload_factor", C.c_uint32), ("pad_", C.c_uint32),
                ("theta", C.c_double), ("selectivity", C.c_double)]


# every symbol include/hjb200.h declares: (restype, argtypes)
SYMBOLS = {
    "hjb_version": (C.c_int, []),
    "hjb_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "hjb_destroy": (C.c_int, [C.c_void_p]),
    "hjb_last_error": (C.c_char_p, [C.c_void_p]),
    "hjb_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hjb_synchronize": (C.c_int, [C.c_void_p]),
    "hjb_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "hjb_kernel_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), u32p, C.c_int]),
    "hjb_kernel_name": (C.c_char_p, [C.c_int]),
    "hjb_npj_device": (C.c_int, [C.c_void_p, C.POINTER(Rel), C.POINTER(Rel), C.POINTER(Opts), C.POINTER(Result)]),
    "hjb_phj_device": (C.c_int, [C.c_void_p, C.POINTER(Rel), C.POINTER(Rel), C.POINTER(Opts), C.POINTER(Result)]),
    "hjb_npj_host": (C.c_int, [C.c_void_p, C.POINTER(Rel), C.POINTER(Rel), C.POINTER(Opts), C.POINTER(Result)]),
    "hjb_phj_host": (C.c_int, [C.c_void_p, C.POINTER(Rel), C.POINTER(Rel), C.POINTER(Opts), C.POINTER(Result)]),
    "hjb_cpra_split": (C.c_int, [C.c_void_p, C.POINTER(Rel), C.POINTER(Rel), C.c_int, C.POINTER(Opts), C.POINTER(Split)]),
    "hjb_cpra_join_local": (C.c_int, [C.c_void_p, C.POINTER(Rel), C.POINTER(Rel), C.c_int, C.c_int, C.POINTER(Opts), C.POINTER(Result)]),
    "hjb_cpra_recv_alloc": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(Recv)]),
    "hjb_ipc_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_ubyte), C.POINTER(C.c_void_p)]),
    "hjb_ipc_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hjb_cpra_count": (C.c_int, [C.c_void_p, C.POINTER(Rel), C.POINTER(Rel), C.c_int, C.POINTER(Opts), u64p, u64p]),
    "hjb_cpra_scatter_peer": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), u64p, u64p, C.POINTER(C.c_float)]),
    "hjb_cpra_bind": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_uint64, C.c_uint64]),
    "hjb_cpra_count_async": (C.c_int, [C.c_void_p, C.POINTER(Rel), C.POINTER(Rel), C.POINTER(Opts), C.c_void_p]),
    "hjb_cpra_scatter_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hjb_cpra_join_async": (C.c_int, [C.c_void_p, C.POINTER(Opts), C.c_uint64, C.c_uint64]),
    "hjb_cpra_sums_dev": (C.c_void_p, [C.c_void_p]),
    "hjb_cpra_finish": (C.c_int, [C.c_void_p, C.POINTER(Result), u64p, u64p]),
    "hjb_cpra_stage_plan": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(Opts), C.POINTER(C.c_int),
                                      C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "hjb_cpra_stage_count_async": (C.c_int, [C.c_void_p, C.POINTER(Rel), C.POINTER(Rel), C.POINTER(Opts), C.c_int, C.c_int, C.c_void_p]),
    "hjb_cpra_stage_scatter_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "hjb_cpra_stage_copy_async": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "hjb_cpra_stage_local_async": (C.c_int, [C.c_void_p, C.POINTER(Opts), C.c_int, C.c_int, C.c_int, C.c_int]),
    "hjb_cpra_split_hot": (C.c_int, [C.c_void_p, C.POINTER(Rel), C.c_void_p, C.c_uint32, C.POINTER(Rel), C.POINTER(Rel)]),
    "hjb_cpra_select_hot": (C.c_int, [C.c_void_p, C.POINTER(Rel), C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, u64p]),
    "hjb_cpra_hot_join": (C.c_int, [C.c_void_p, C.POINTER(Rel), C.POINTER(Rel)]),
    "hjb_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "hjb_host_unregister": (C.c_int, [C.c_void_p]),
    "hjb_cpra_count_async_host": (C.c_int, [C.c_void_p, C.POINTER(Rel), C.POINTER(Rel), C.POINTER(Opts), C.c_void_p]),
    "hjb_cpra_finish_host": (C.c_int, [C.c_void_p, C.POINTER(Result), u64p, u64p]),
    "hjb_hash_factor": (C.c_uint32, [C.c_uint32, C.c_int]),
    "hjb_histogram": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, u32p, C.c_uint32, C.c_int, C.c_int]),
    "hjb_partition_pass": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, u32p, C.c_void_p, C.c_void_p, u32p, C.c_uint32, C.c_int, C.c_int]),
    "hjb_npj_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32]),
    "hjb_relation_write": (C.c_int, [C.c_char_p, C.c_int, C.c_uint64, u32p, u32p]),
    "hjb_relation_read": (C.c_int, [C.c_char_p, C.c_int, C.c_uint64, u32p, u32p]),
    "hjb_generate": (C.c_int, [C.c_void_p, C.POINTER(Gen), C.c_void_p, C.c_void_p]),
    "hjb_column_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, u64p]),
    "hjb_rows_fingerprint": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, u64p]),
}

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -m hash_join_codes_knl_b200.build` "
                              "(nvcc, sm_100a). hash_join_codes_knl_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)      # AttributeError here = the library does not match the header
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib
