"""ctypes binding of include/nanomod_b200.h (the C ABI in nanomod_b200/_C/libnanomod_b200.so).

There is no fallback: if the shared library has not been built (``python -m nanomod_b200.build``
or ``__graft_entry__.build()``) or no sm_100 GPU is present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NANOMOD_B200_LIB selects an experimental build variant of the same library (see build.py)
LIB_PATH = os.environ.get("NANOMOD_B200_LIB") or os.path.join(_HERE, "_C", "libnanomod_b200.so")

NM_OK = 0
NM_COMBINE_NONE, NM_COMBINE_FISHER, NM_COMBINE_STOUFFER = 0, 1, 2
NM_MAX_NB = 32
NM_LANE_TIER_MAX = 128
NM_DS_MAX_READS = 256
NM_DS_DEEP_MAX_READS = 32768
NM_DS_DEEP_MAX_COV = 1024
NM_DS_MAX_TIMES = 1024
NM_RECORD_BYTES = 28

NM_ERR_BAD_ARG = 1
ERROR_NAMES = {1: "NM_ERR_BAD_ARG", 2: "NM_ERR_BAD_PARAM", 3: "NM_ERR_CUDA", 4: "NM_ERR_OOM",
               5: "NM_ERR_TOO_DEEP", 6: "NM_ERR_NO_DEVICE"}

# every symbol include/nanomod_b200.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = ["nm_version", "nm_padded_len", "nm_create", "nm_destroy", "nm_last_error",
                    "nm_detect_device", "nm_detect_device_async", "nm_detect_finish", "nm_detect_host", "nm_launch_count", "nm_last_timings", "nm_last_path", "nm_last_grid_tiles", "nm_grid_selftest", "nm_rank_head_device", "nm_rank_head_select_device", "nm_arm_head_select", "nm_head_fired", "nm_head_set_peers", "nm_peer_alloc", "nm_peer_open", "nm_peer_close", "nm_peer_free",
                    "nm_set_sm_limit", "nm_sm_count", "nm_rank_device", "nm_rank_host", "nm_pack_records_device",
                    "nm_format_bound", "nm_format_sign_test"]


class NmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("%s (%d): %s" % (ERROR_NAMES.get(code, "NM_ERR_?"), code, msg))
        self.code = code


class nm_params(C.Structure):
    _fields_ = [("min_coverage", C.c_int32), ("nb", C.c_int32), ("weights_dif", C.c_double),
                ("combine", C.c_int32), ("want_u", C.c_int32), ("want_t", C.c_int32),
                ("reserved", C.c_int32), ("ds_times", C.c_int32), ("ds_index", C.c_int32),
                ("ds_seed", C.c_uint64)]


class nm_pileup(C.Structure):
    _fields_ = [("vals0", C.c_void_p), ("off0", C.c_void_p), ("vals1", C.c_void_p),
                ("off1", C.c_void_p), ("pos", C.c_void_p), ("seg", C.c_void_p),
                ("n_pos", C.c_int64), ("seg_cov", C.c_void_p), ("n_seg", C.c_int64),
                ("vals0_i16", C.c_void_p), ("vals1_i16", C.c_void_p), ("i16_unit", C.c_double),
                ("i16_total0", C.c_int64), ("i16_total1", C.c_int64)]


TABLE_FIELDS = ["row_pos_index", "n0", "n1", "ks_dnum", "ks_d", "ks_p", "two_u", "u_stat", "u_p",
                "t_stat", "t_p", "fisher_stat", "fisher_p", "stouffer_stat", "stouffer_p", "flags",
                "moments"]
TABLE_DTYPES = {"row_pos_index": "int32", "n0": "int32", "n1": "int32", "ks_dnum": "int32",
                "ks_d": "float64", "ks_p": "float64", "two_u": "int64", "u_stat": "float64",
                "u_p": "float64", "t_stat": "float64", "t_p": "float64", "fisher_stat": "float64",
                "fisher_p": "float64", "stouffer_stat": "float64", "stouffer_p": "float64",
                "flags": "uint8", "moments": "float64"}
TABLE_WIDTH = {"moments": 4}  # columns that hold several values per row


class nm_head_geometry(C.Structure):
    _fields_ = [("row_pos_index", C.c_void_p), ("pos", C.c_void_p), ("seg", C.c_void_p), ("row_offset", C.c_int64),
                ("n_rows_total", C.c_int64), ("nearby", C.c_int32), ("reserved", C.c_int32)]


HEAD_ROW_DTYPE = [("row", "<i8"), ("seg", "<i4"), ("pos", "<i4"), ("full_nbhd", "<i4"), ("reserved", "<i4"),
                  ("key", "<u8", (3,))]


NM_MAX_PEERS = 16
NM_IPC_HANDLE_BYTES = 64


class nm_head_peers(C.Structure):
    _fields_ = [("n_peers", C.c_int32), ("epoch", C.c_int32), ("base", C.c_void_p * NM_MAX_PEERS)]


class nm_text_columns(C.Structure):
    _fields_ = [("seg_chrom", C.POINTER(C.c_char_p)), ("seg_strand", C.POINTER(C.c_char_p)), ("n_seg", C.c_int32),
                ("reserved", C.c_int32), ("n_rows", C.c_int64)] + \
               [(n, C.c_void_p) for n in ("seg", "pos", "base", "n0", "n1", "u_stat", "u_p", "t_stat", "t_p", "ks_d",
                                          "ks_p", "comb_stat", "comb_p")]


class nm_table(C.Structure):
    _fields_ = [(name, C.c_void_p) for name in TABLE_FIELDS]


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "nanomod_b200: %s is missing -- build it with `python -m nanomod_b200.build` "
            "(needs nvcc).  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.nm_version.restype = C.c_int
    lib.nm_padded_len.restype = C.c_int64
    lib.nm_padded_len.argtypes = [C.c_int64]
    lib.nm_create.restype = C.c_int
    lib.nm_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.nm_destroy.restype = None
    lib.nm_destroy.argtypes = [C.c_void_p]
    lib.nm_last_error.restype = C.c_char_p
    lib.nm_last_error.argtypes = [C.c_void_p]
    lib.nm_launch_count.restype = C.c_int64
    lib.nm_launch_count.argtypes = [C.c_void_p]
    lib.nm_set_sm_limit.restype = C.c_int
    lib.nm_set_sm_limit.argtypes = [C.c_void_p, C.c_int]
    lib.nm_sm_count.restype = C.c_int
    lib.nm_sm_count.argtypes = [C.c_void_p]
    lib.nm_last_path.restype = C.c_int
    lib.nm_last_path.argtypes = [C.c_void_p]
    lib.nm_grid_selftest.restype = C.c_int
    lib.nm_grid_selftest.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.nm_last_grid_tiles.restype = C.c_int64
    lib.nm_last_grid_tiles.argtypes = [C.c_void_p]
    lib.nm_last_timings.restype = C.c_int
    lib.nm_last_timings.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.nm_detect_device.restype = C.c_int
    lib.nm_detect_device.argtypes = [C.c_void_p, C.POINTER(nm_pileup), C.POINTER(nm_params),
                                     C.POINTER(nm_table), C.POINTER(C.c_int64), C.c_void_p]
    lib.nm_detect_host.restype = C.c_int
    lib.nm_detect_host.argtypes = [C.c_void_p, C.POINTER(nm_pileup), C.POINTER(nm_params),
                                   C.POINTER(nm_table), C.POINTER(C.c_int64)]
    lib.nm_rank_device.restype = C.c_int
    lib.nm_rank_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                   C.c_void_p, C.c_void_p]
    lib.nm_rank_host.restype = C.c_int
    lib.nm_rank_head_device.restype = C.c_int
    lib.nm_rank_head_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64,
                                        C.POINTER(nm_head_geometry), C.c_void_p, C.c_int64, C.POINTER(C.c_int64),
                                        C.c_void_p]
    lib.nm_detect_device_async.restype = C.c_int
    lib.nm_detect_device_async.argtypes = [C.c_void_p, C.POINTER(nm_pileup), C.POINTER(nm_params), C.POINTER(nm_table),
                                           C.c_void_p, C.POINTER(C.c_int)]
    lib.nm_detect_finish.restype = C.c_int
    lib.nm_detect_finish.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int)]
    lib.nm_head_set_peers.restype = C.c_int
    lib.nm_head_set_peers.argtypes = [C.c_void_p, C.POINTER(nm_head_peers)]
    lib.nm_peer_alloc.restype = C.c_int
    lib.nm_peer_alloc.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.c_char_p]
    lib.nm_peer_open.restype = C.c_int
    lib.nm_peer_open.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.nm_peer_close.restype = C.c_int
    lib.nm_peer_close.argtypes = [C.c_void_p, C.c_void_p]
    lib.nm_peer_free.restype = C.c_int
    lib.nm_peer_free.argtypes = [C.c_void_p, C.c_void_p]
    lib.nm_arm_head_select.restype = C.c_int
    lib.nm_arm_head_select.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64,
                                       C.POINTER(nm_head_geometry), C.c_void_p, C.c_int64]
    lib.nm_head_fired.restype = C.c_int
    lib.nm_head_fired.argtypes = [C.c_void_p]
    lib.nm_rank_head_select_device.restype = C.c_int
    lib.nm_rank_head_select_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64,
                                               C.POINTER(nm_head_geometry), C.c_void_p, C.c_int64, C.c_void_p]
    lib.nm_rank_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                 C.c_void_p]
    lib.nm_format_bound.restype = C.c_int64
    lib.nm_format_bound.argtypes = [C.POINTER(nm_text_columns)]
    lib.nm_format_sign_test.restype = C.c_int64
    lib.nm_format_sign_test.argtypes = [C.POINTER(nm_text_columns), C.c_int, C.c_void_p, C.c_int64]
    lib.nm_pack_records_device.restype = C.c_int
    lib.nm_pack_records_device.argtypes = [C.c_void_p, C.POINTER(nm_table), C.c_int64, C.c_int64, C.c_int,
                                           C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


def padded_len(nvals: int) -> int:
    """Floats a vals buffer must have allocated for nvals values (nm_padded_len)."""
    return (int(nvals) + 3) // 4 * 4 + 4


class Handle:
    """Owns one nm_handle (one GPU).  Not thread-safe; use one per thread / rank."""

    def __init__(self, device: int = 0):
        self._lib = load()
        self._h = C.c_void_p()
        rc = self._lib.nm_create(int(device), C.byref(self._h))
        if rc != NM_OK:
            raise NmError(rc, self._lib.nm_last_error(None).decode())
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.nm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != NM_OK:
            raise NmError(rc, self._lib.nm_last_error(self._h).decode())

    @property
    def launch_count(self) -> int:
        return int(self._lib.nm_launch_count(self._h))

    @property
    def sm_count(self) -> int:
        return int(self._lib.nm_sm_count(self._h))

    def set_sm_limit(self, n_sms: int) -> None:
        """Let the persistent lane kernel use only n_sms SMs (0 = all): leaves room for NCCL."""
        self._check(self._lib.nm_set_sm_limit(self._h, int(n_sms)))

    def last_path(self) -> int:
        """0 general, 1 dense, 2 dense (speculative launch), 3 / 4 speculative launch refused and the call re-run
        dense / on the general path"""
        return int(self._lib.nm_last_path(self._h))

    def last_grid_tiles(self) -> int:
        """32-position tiles of the last call that the lane tier sorted as packed 16-bit grid keys"""
        return int(self._lib.nm_last_grid_tiles(self._h))

    def grid_selftest(self):
        """(violations, passes) of the grid-key check over all 2^32 float32 patterns, on the device"""
        v, n = C.c_int64(), C.c_int64()
        self._check(self._lib.nm_grid_selftest(self._h, C.byref(v), C.byref(n)))
        return int(v.value), int(n.value)

    def last_timings(self):
        """Device ms of the last call: {'plan','lane','deep','combine'} (CUDA events)."""
        ms = (C.c_double * 4)()
        self._check(self._lib.nm_last_timings(self._h, ms))
        return {"plan": ms[0], "lane": ms[1], "deep": ms[2], "combine": ms[3]}

    def detect_host(self, pileup: nm_pileup, params: nm_params, table: nm_table) -> int:
        n_rows = C.c_int64(0)
        self._check(self._lib.nm_detect_host(self._h, C.byref(pileup), C.byref(params),
                                             C.byref(table), C.byref(n_rows)))
        return int(n_rows.value)

    def detect_device(self, pileup: nm_pileup, params: nm_params, table: nm_table,
                      stream: int = 0) -> int:
        n_rows = C.c_int64(0)
        self._check(self._lib.nm_detect_device(self._h, C.byref(pileup), C.byref(params),
                                               C.byref(table), C.byref(n_rows),
                                               C.c_void_p(int(stream))))
        return int(n_rows.value)

    def detect_device_async(self, pileup: nm_pileup, params: nm_params, table: nm_table, stream: int = 0) -> int:
        """nm_detect_device_async: queues the call, returns its ticket (finish with ``detect_finish``)."""
        ticket = C.c_int(-1)
        self._check(self._lib.nm_detect_device_async(self._h, C.byref(pileup), C.byref(params), C.byref(table),
                                                     C.c_void_p(int(stream)), C.byref(ticket)))
        return int(ticket.value)

    def detect_finish(self, ticket: int):
        """nm_detect_finish -> (n_rows, head_fired)"""
        n_rows = C.c_int64(0)
        fired = C.c_int(0)
        self._check(self._lib.nm_detect_finish(self._h, int(ticket), C.byref(n_rows), C.byref(fired)))
        return int(n_rows.value), bool(fired.value)

    def rank_host(self, key_comb, key_ks, key_u, n_rows: int, reverse: bool, order) -> None:
        """nm_rank_host on raw host addresses (0 / None = key absent)."""
        self._check(self._lib.nm_rank_host(self._h, C.c_void_p(key_comb or 0), C.c_void_p(key_ks),
                                           C.c_void_p(key_u or 0), int(n_rows), int(bool(reverse)),
                                           C.c_void_p(order)))

    def rank_device(self, key_comb, key_ks, key_u, n_rows: int, reverse: bool, order, stream: int = 0) -> None:
        self._check(self._lib.nm_rank_device(self._h, C.c_void_p(key_comb or 0), C.c_void_p(key_ks),
                                             C.c_void_p(key_u or 0), int(n_rows), int(bool(reverse)),
                                             C.c_void_p(order), C.c_void_p(int(stream))))

    def rank_head_device(self, key_comb, key_ks, key_u, n_rows: int, reverse: bool, want: int, geometry, rows_out,
                         cap: int, stream: int = 0) -> int:
        n_head = C.c_int64(0)
        self._check(self._lib.nm_rank_head_device(self._h, key_comb, key_ks, key_u, n_rows, 1 if reverse else 0, want,
                                                  None if geometry is None else C.byref(geometry), rows_out, cap,
                                                  C.byref(n_head), C.c_void_p(stream)))
        return int(n_head.value)

    def rank_head_select_device(self, key_comb, key_ks, key_u, n_rows: int, reverse: bool, want: int, geometry,
                                records_dev: int, cap: int, stream: int = 0) -> None:
        self._check(self._lib.nm_rank_head_select_device(self._h, key_comb, key_ks, key_u, n_rows, 1 if reverse else 0, want,
                                                         None if geometry is None else C.byref(geometry), records_dev,
                                                         cap, C.c_void_p(stream)))

    def arm_head_select(self, key_comb, key_ks, key_u, n_rows: int, reverse: bool, want: int, geometry,
                        records_dev: int, cap: int) -> None:
        """nm_arm_head_select: the next detect_device call launches this selection before its host wait"""
        self._check(self._lib.nm_arm_head_select(self._h, key_comb, key_ks, key_u, n_rows, 1 if reverse else 0, want,
                                                 None if geometry is None else C.byref(geometry), records_dev, cap))

    def head_fired(self) -> bool:
        return bool(self._lib.nm_head_fired(self._h))

    def head_set_peers(self, bases, epoch: int) -> None:
        """nm_head_set_peers: the next head selection also stores its header and records into these device
        addresses (this rank's section of each peer's gathered-heads buffer); ``bases`` empty/None clears."""
        if not bases:
            self._check(self._lib.nm_head_set_peers(self._h, None))
            return
        pr = nm_head_peers()
        pr.n_peers = len(bases)
        pr.epoch = int(epoch)
        for i, b in enumerate(bases):
            pr.base[i] = int(b)
        self._check(self._lib.nm_head_set_peers(self._h, C.byref(pr)))

    def peer_alloc(self, nbytes: int):
        """nm_peer_alloc -> (device address, 64-byte CUDA IPC handle)"""
        ptr = C.c_void_p(0)
        hd = C.create_string_buffer(NM_IPC_HANDLE_BYTES)
        self._check(self._lib.nm_peer_alloc(self._h, int(nbytes), C.byref(ptr), hd))
        return int(ptr.value), bytes(hd.raw)

    def peer_open(self, ipc_handle: bytes) -> int:
        ptr = C.c_void_p(0)
        self._check(self._lib.nm_peer_open(self._h, C.create_string_buffer(bytes(ipc_handle), NM_IPC_HANDLE_BYTES), C.byref(ptr)))
        return int(ptr.value)

    def peer_close(self, ptr: int) -> None:
        self._check(self._lib.nm_peer_close(self._h, C.c_void_p(int(ptr))))

    def peer_free(self, ptr: int) -> None:
        self._check(self._lib.nm_peer_free(self._h, C.c_void_p(int(ptr))))

    def pack_records_device(self, table: nm_table, row_lo: int, n: int, which_combine: int, records: int,
                            stream: int = 0) -> None:
        """nm_pack_records_device: 28-byte result records of rows [row_lo, row_lo + n)."""
        self._check(self._lib.nm_pack_records_device(self._h, C.byref(table), int(row_lo), int(n), int(which_combine),
                                                     C.c_void_p(records), C.c_void_p(int(stream))))


def format_sign_test(seg_names, seg, pos, base, n0, n1, u_stat, u_p, t_stat, t_p, ks_d, ks_p, comb_stat, comb_p,
                     n_threads: int = 0):
    """nm_format_sign_test: the text of ``<FileID>_sign_test.txt`` for these columns (numpy arrays;
    optional ones may be None) as a uint8 numpy array.  Host code only -- needs the library, not a
    GPU."""
    import numpy as np
    lib = load()
    n = int(len(pos))
    if n == 0:
        return np.zeros(0, np.uint8)
    keep = []

    def arr(x, dt):
        if x is None:
            return None
        a = np.ascontiguousarray(x, dtype=dt)
        keep.append(a)
        return a.ctypes.data
    chrom = (C.c_char_p * len(seg_names))(*[str(k[0]).encode() for k in seg_names])
    strand = (C.c_char_p * len(seg_names))(*[str(k[1]).encode() for k in seg_names])
    cols = nm_text_columns(chrom, strand, len(seg_names), 0, n, arr(seg, np.int32), arr(pos, np.int32),
                           arr(base, np.uint8), arr(n0, np.int32), arr(n1, np.int32), arr(u_stat, np.float64),
                           arr(u_p, np.float64), arr(t_stat, np.float64), arr(t_p, np.float64),
                           arr(ks_d, np.float64), arr(ks_p, np.float64), arr(comb_stat, np.float64),
                           arr(comb_p, np.float64))
    if n_threads <= 0:
        n_threads = min(32, os.cpu_count() or 1)
    cap = n * 160 + int(lib.nm_format_bound(C.byref(cols)))
    while True:
        buf = np.empty(cap, np.uint8)
        got = int(lib.nm_format_sign_test(C.byref(cols), int(n_threads), C.c_void_p(buf.ctypes.data), cap))
        if got >= 0:
            return buf[:got]
        if got == -1:
            raise NmError(1, "nm_format_sign_test: bad argument")
        cap = -got
