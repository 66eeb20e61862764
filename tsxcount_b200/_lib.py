"""ctypes binding of libtsxcuda.so (include/tsxcount_cuda.h).

The library is built in-tree by `make lib` / `__graft_entry__.build()` into tsxcount_b200/lib/.
There is no fallback of any kind: a missing library raises, and every compute entry point of the
library itself fails with TSXC_E_CUDA when no sm_100 device is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TSXC_LIB") or os.path.join(_HERE, "lib", "libtsxcuda.so")   # TSXC_LIB: A/B kernel builds

TSXC_OK = 0
TSXC_E_INVALID = 1
TSXC_E_CUDA = 2
TSXC_E_NOMEM = 3
TSXC_E_UNSUPPORTED = 4
TSXC_E_COUNT_SATURATED = 5
TSXC_E_IO = 6
TSXC_E_TABLE_FULL = 42

TSXC_FLAG_NONE = 0
TSXC_FLAG_EXACT_S = 1
TSXC_FLAG_NO_WARP_AGG = 2
TSXC_FLAG_DIRECT = 4
TSXC_FLAG_CANONICAL = 8

TSXC_IPC_HANDLE_BYTES = 64


class TsxcStats(C.Structure):
    _fields_ = [
        ("k", C.c_uint32), ("l", C.c_uint32), ("s", C.c_uint32),
        ("key_words", C.c_uint32), ("entry_words", C.c_uint32), ("value_bits", C.c_uint32),
        ("quotient_bits", C.c_uint32), ("reprobe_bits", C.c_uint32), ("slots_per_bucket", C.c_uint32),
        ("n_shards", C.c_uint32), ("shard_rank", C.c_uint32), ("reserved0", C.c_uint32),
        ("n_slots", C.c_uint64), ("table_bytes", C.c_uint64), ("distinct", C.c_uint64),
        ("overflow_entries", C.c_uint64), ("used_slots", C.c_uint64), ("kmers_added", C.c_uint64),
        ("max_reprobe", C.c_uint64), ("error_flags", C.c_uint64),
        ("kernel_launches", C.c_uint64), ("main_kernel_launches", C.c_uint64), ("main_kernel_ms", C.c_double),
        ("partition_ms", C.c_double), ("insert_ms", C.c_double),
        ("hist_ms", C.c_double), ("part1_ms", C.c_double), ("part2_ms", C.c_double),
        ("chunk_cap_keys", C.c_uint64), ("group_cap_keys", C.c_uint64),
        ("radix_digit1_bits", C.c_uint32), ("radix_digit2_bits", C.c_uint32),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class TsxcGenParams(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("n_reads", C.c_uint64), ("read_len", C.c_uint32), ("mode", C.c_uint32),
        ("genome_len", C.c_uint64), ("sub_rate_q16", C.c_uint32), ("reserved", C.c_uint32),
    ]


class TsxcRouteInfo(C.Structure):
    _fields_ = [
        ("n_shards", C.c_uint32), ("shard_rank", C.c_uint32), ("bins", C.c_uint32), ("bins_per_shard", C.c_uint32),
        ("key_words", C.c_uint32), ("reserved", C.c_uint32), ("recv_cap_keys", C.c_uint64),
    ]


_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/tsxcount_cuda.h declares
PROTOTYPES = {
    "tsxc_key_words": (C.c_uint32, [C.c_uint32]),
    "tsxc_abi_version": (C.c_int, []),
    "tsxc_device_count": (C.c_int, []),
    "tsxc_status_string": (C.c_char_p, [C.c_int]),
    "tsxc_last_error": (C.c_char_p, [_vp]),
    "tsxc_create": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.POINTER(_vp)]),
    "tsxc_create_shard": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32,
                                    C.POINTER(_vp)]),
    "tsxc_destroy": (C.c_int, [_vp]),
    "tsxc_clear": (C.c_int, [_vp]),
    "tsxc_trim": (C.c_int, [_vp]),
    "tsxc_stream": (_vp, [_vp]),
    "tsxc_mark": (C.c_int, [_vp, C.c_int]),
    "tsxc_mark_elapsed_ms": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "tsxc_add_reads": (C.c_int, [_vp, _vp, _vp, C.c_uint64]),
    "tsxc_add_reads_device": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint64]),
    "tsxc_add_kmers": (C.c_int, [_vp, _vp, C.c_uint64]),
    "tsxc_add_kmers_device": (C.c_int, [_vp, _vp, C.c_uint64]),
    "tsxc_sync": (C.c_int, [_vp]),
    "tsxc_lookup": (C.c_int, [_vp, _vp, C.c_uint64, _vp]),
    "tsxc_lookup_device": (C.c_int, [_vp, _vp, C.c_uint64, _vp]),
    "tsxc_distinct": (C.c_int, [_vp, _u64p]),
    "tsxc_dump": (C.c_int, [_vp, _vp, _vp, C.c_uint64, _u64p]),
    "tsxc_dump_file": (C.c_int, [_vp, C.c_char_p]),
    "tsxc_stats": (C.c_int, [_vp, C.POINTER(TsxcStats)]),
    "tsxc_histogram": (C.c_int, [_vp, _vp, C.c_uint32]),
    "tsxc_route_info": (C.c_int, [_vp, C.POINTER(TsxcRouteInfo)]),
    "tsxc_route_recv_buffer": (C.c_int, [_vp, C.c_uint64, C.POINTER(_vp), _u64p]),
    "tsxc_route_set_peers": (C.c_int, [_vp, C.POINTER(_vp), C.c_uint64]),
    "tsxc_route_begin": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32)]),
    "tsxc_route_hist": (C.c_int, [_vp, C.c_uint32, _vp]),
    "tsxc_route_send": (C.c_int, [_vp, C.c_uint32, _vp]),
    "tsxc_route_insert": (C.c_int, [_vp]),
    "tsxc_pack_reads": (C.c_int, [_vp, _vp, C.c_uint64, _vp, _vp, C.c_uint64, _u64p, _u64p]),
    "tsxc_gen_reads_device": (C.c_int, [C.POINTER(TsxcGenParams), C.c_uint64, C.c_uint64, C.c_int, _vp, _vp, _vp]),
    "tsxc_k0_random_rmw": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_float)]),
    "tsxc_k0_windowed": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_float)]),
    "tsxc_k0_region_sweep": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(C.c_float)]),
    "tsxc_host_alloc": (C.c_int, [C.c_uint64, C.POINTER(_vp)]),
    "tsxc_host_free": (C.c_int, [_vp]),
    "tsxc_device_alloc": (C.c_int, [C.c_int, C.c_uint64, C.POINTER(_vp)]),
    "tsxc_device_free": (C.c_int, [C.c_int, _vp]),
    "tsxc_memcpy": (C.c_int, [C.c_int, _vp, _vp, C.c_uint64, C.c_int]),
    "tsxc_enable_peer_access": (C.c_int, [C.c_int, C.c_int]),
    "tsxc_ipc_export_mem": (C.c_int, [C.c_int, _vp, _vp]),
    "tsxc_ipc_open_mem": (C.c_int, [C.c_int, _vp, C.POINTER(_vp)]),
    "tsxc_ipc_close_mem": (C.c_int, [C.c_int, _vp]),
    "tsxc_ipc_event_create": (C.c_int, [C.c_int, C.POINTER(_vp), _vp]),
    "tsxc_ipc_event_open": (C.c_int, [C.c_int, _vp, C.POINTER(_vp)]),
    "tsxc_event_destroy": (C.c_int, [C.c_int, _vp]),
    "tsxc_event_record": (C.c_int, [C.c_int, _vp, _vp]),
    "tsxc_stream_wait_event": (C.c_int, [C.c_int, _vp, _vp]),
    "tsxc_copy_async": (C.c_int, [C.c_int, _vp, _vp, C.c_uint64, _vp]),
    "tsxc_debug_hash": (C.c_int, [C.c_uint32, _vp, _vp]),
    "tsxc_debug_unhash": (C.c_int, [C.c_uint32, _vp, _vp]),
    "tsxc_debug_canonical": (C.c_int, [C.c_uint32, _vp, _vp]),
    "tsxc_debug_sparse_round": (C.c_int, [C.c_uint32, _vp, _vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _vp, _vp]),
    "tsxc_debug_layout": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(TsxcStats)]),
}

_lib = None


def load():
    """Load libtsxcuda.so and declare every prototype.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make lib` (or __graft_entry__.build()). "
            "tsxcount_b200 has no CPU or pure-Python fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class TsxcError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"tsxc status {status}: {message}")
        self.status = status


def check(status, handle=None):
    if status != TSXC_OK:
        lib = load()
        msg = lib.tsxc_last_error(handle) or b""
        raise TsxcError(status, (lib.tsxc_status_string(status) or b"").decode() + (": " + msg.decode() if msg else ""))
