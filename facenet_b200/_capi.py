"""ctypes binding of the C ABI in include/facenet_b200.h.

Arrays cross the boundary as DLPack tensors: any object with ``__dlpack__`` (NumPy,
PyTorch, TensorFlow) is exported to a capsule and the ``DLTensor`` inside is BORROWED for
the duration of the call (the capsule keeps ownership, nothing is copied on this side).
NumPy arrays arrive as kDLCPU and are staged to the GPU by the library; GPU tensors
(kDLCUDA) are used in place.

There is no CPU fallback: if the shared library is missing or no CUDA device is present
the import / handle creation raises.
"""
import ctypes
import os
import threading
from pathlib import Path

import numpy as np

_LIB_PATH = Path(__file__).resolve().parent / '_lib' / 'libfacenet_b200.so'

FNB_OK, FNB_ERR_INVALID, FNB_ERR_NOT_NORMALIZED, FNB_ERR_BAD_METRIC, FNB_ERR_CUDA, FNB_ERR_UNSUPPORTED = range(6)
MODES = {'fp16x3': 0, 'tf32x3': 1, 'tf32': 2, 'bf16': 3, 'fp16': 4, 'fp16f8': 5, 'auto': 6}
MODE_NAMES = {v: k for k, v in MODES.items()}
MAX_THRESHOLDS = 127

EXPORTS = ('fnb_version', 'fnb_default_options', 'fnb_create', 'fnb_destroy', 'fnb_last_error', 'fnb_device_info', 'fnb_set_stream',
           'fnb_pairwise', 'fnb_pair_histogram_bins', 'fnb_counts_from_bins', 'fnb_pair_histogram',
           'fnb_region_histogram_bins', 'fnb_confidence_from_last_bins', 'fnb_mine', 'fnb_mine_batched', 'fnb_mine_check',
           'fnb_mine_select_kth', 'fnb_false_pairs', 'fnb_pair_cross_entropy', 'fnb_logits_cross_entropy',
           'fnb_comm_unique_id', 'fnb_comm_init', 'fnb_comm_destroy', 'fnb_comm_info', 'fnb_comm_shared_queue', 'fnb_comm_last_error',
           'fnb_pair_histogram_sharded', 'fnb_debug_chunk_plan')


class DLDevice(ctypes.Structure):
    _fields_ = [('device_type', ctypes.c_int32), ('device_id', ctypes.c_int32)]


class DLDataType(ctypes.Structure):
    _fields_ = [('code', ctypes.c_uint8), ('bits', ctypes.c_uint8), ('lanes', ctypes.c_uint16)]


class DLTensor(ctypes.Structure):
    _fields_ = [('data', ctypes.c_void_p), ('device', DLDevice), ('ndim', ctypes.c_int32), ('dtype', DLDataType),
                ('shape', ctypes.POINTER(ctypes.c_int64)), ('strides', ctypes.POINTER(ctypes.c_int64)),
                ('byte_offset', ctypes.c_uint64)]


class Options(ctypes.Structure):
    _fields_ = [('mode', ctypes.c_int32), ('metric', ctypes.c_int32), ('atol', ctypes.c_float), ('eps', ctypes.c_float),
                ('rank', ctypes.c_int32), ('world', ctypes.c_int32), ('cta_group', ctypes.c_int32),
                ('region_rows', ctypes.c_int32), ('cuts', ctypes.POINTER(ctypes.c_float)),
                ('max_ctas', ctypes.c_int32), ('force_checked', ctypes.c_int32), ('debug', ctypes.c_int32),
                ('cluster_pairs', ctypes.c_int32), ('normalize', ctypes.c_int32), ('theta', ctypes.c_float),
                ('raw_distance', ctypes.c_int32), ('subset_rows', ctypes.c_int32), ('shard_mod', ctypes.c_int32),
                ('shard_lo', ctypes.c_int32), ('shard_width', ctypes.c_int32), ('shard_slots', ctypes.POINTER(ctypes.c_int32)),
                ('panel_window', ctypes.c_int32), ('strict_tiles', ctypes.c_int32), ('bias_correction', ctypes.c_int32),
                ('streamed', ctypes.c_int32), ('tile_queue', ctypes.c_int32)]


class Stats(ctypes.Structure):
    _fields_ = [('n_pairs', ctypes.c_uint64), ('eps_window', ctypes.c_uint64), ('smin', ctypes.c_float),
                ('smax', ctypes.c_float), ('max_abs', ctypes.c_float), ('kernel_ms', ctypes.c_float),
                ('prepare_ms', ctypes.c_float), ('tiles', ctypes.c_uint64), ('kernel_launches', ctypes.c_uint32),
                ('eps_counted', ctypes.c_float), ('grid_ctas', ctypes.c_uint32), ('mode_used', ctypes.c_int32), ('peakedness', ctypes.c_float),
                ('panel_window', ctypes.c_int32), ('error_bound', ctypes.c_float), ('fallback', ctypes.c_int32), ('h2d_ms', ctypes.c_float), ('h2d_bytes', ctypes.c_uint64),
                ('gather_ms', ctypes.c_float), ('streamed_chunks', ctypes.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != 'reserved'}


class Region(ctypes.Structure):
    _fields_ = [('row_begin', ctypes.c_int32), ('row_end', ctypes.c_int32), ('col_begin', ctypes.c_int32),
                ('col_end', ctypes.c_int32), ('tri', ctypes.c_int32), ('key', ctypes.c_int32)]


REGION_DTYPE = np.dtype([('row_begin', '<i4'), ('row_end', '<i4'), ('col_begin', '<i4'), ('col_end', '<i4'),
                         ('tri', '<i4'), ('key', '<i4')])

_lib = None
_lib_lock = threading.Lock()


def load_library():
    """Load libfacenet_b200.so (built by ``python -m facenet_b200.build`` / ``__graft_entry__.build()``)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not _LIB_PATH.exists():
            raise ImportError('%s is missing: run `python -m facenet_b200.build` (needs nvcc). '
                              'facenet_b200 has no CPU fallback.' % _LIB_PATH)
        lib = ctypes.CDLL(str(_LIB_PATH))
        c = ctypes
        P = c.POINTER
        lib.fnb_version.restype = c.c_int
        lib.fnb_default_options.argtypes = [P(Options)]
        lib.fnb_default_options.restype = None
        lib.fnb_create.argtypes = [c.c_int, P(c.c_void_p)]
        lib.fnb_destroy.argtypes = [c.c_void_p]
        lib.fnb_destroy.restype = None
        lib.fnb_last_error.argtypes = [c.c_void_p]
        lib.fnb_last_error.restype = c.c_char_p
        lib.fnb_device_info.argtypes = [c.c_void_p, P(c.c_int), P(c.c_int), P(c.c_int), P(c.c_uint64)]
        lib.fnb_set_stream.argtypes = [c.c_void_p, c.c_void_p]
        lib.fnb_pairwise.argtypes = [c.c_void_p, P(DLTensor), P(DLTensor), P(Options), P(DLTensor), P(c.c_float)]
        lib.fnb_pair_histogram_bins.argtypes = [c.c_void_p, P(DLTensor), P(DLTensor), P(c.c_double), c.c_int, P(Options),
                                                P(DLTensor), P(Stats)]
        lib.fnb_counts_from_bins.argtypes = [P(c.c_double), c.c_int, P(Options), P(c.c_uint64), P(c.c_uint64),
                                             P(c.c_uint64), P(c.c_uint64), P(c.c_uint64)]
        lib.fnb_pair_histogram.argtypes = [c.c_void_p, P(DLTensor), P(DLTensor), P(c.c_double), c.c_int, P(Options),
                                           P(c.c_uint64), P(c.c_uint64), P(c.c_uint64), P(c.c_uint64), P(Stats)]
        lib.fnb_region_histogram_bins.argtypes = [c.c_void_p, P(DLTensor), P(c.c_int64), P(c.c_int32), P(Region), c.c_int,
                                                  c.c_int, P(c.c_double), c.c_int, P(Options), P(c.c_uint64), P(Stats)]
        lib.fnb_confidence_from_last_bins.argtypes = [c.c_void_p, c.c_int, P(c.c_double), P(c.c_double), P(c.c_double),
                                                      c.c_int, P(Options), c.c_double, P(c.c_double), P(c.c_double),
                                                      P(c.c_double), P(c.c_double), P(c.c_int32), P(c.c_double)]
        lib.fnb_mine.argtypes = [c.c_void_p, P(DLTensor), P(DLTensor), c.c_float, P(Options), P(c.c_int32), P(c.c_int32),
                                 c.c_int, P(c.c_int32), P(c.c_int32), P(c.c_int32), P(Stats)]
        lib.fnb_mine_batched.argtypes = [c.c_void_p, P(DLTensor), P(DLTensor), c.c_int, c.c_float, P(Options), c.c_int, P(DLTensor),
                                         P(DLTensor), P(DLTensor), P(DLTensor), P(DLTensor), P(DLTensor), P(Stats)]
        lib.fnb_mine_check.argtypes = [c.c_void_p, P(c.c_int32), P(Stats)]
        lib.fnb_mine_select_kth.argtypes = [c.c_void_p, P(DLTensor), P(DLTensor), P(DLTensor), c.c_float, P(DLTensor)]
        lib.fnb_false_pairs.argtypes = [c.c_void_p, P(DLTensor), P(DLTensor), c.c_double, P(Options), c.c_longlong, P(c.c_int32),
                                        P(c.c_int32), P(c.c_float), P(c.c_uint64), P(Stats)]
        lib.fnb_pair_cross_entropy.argtypes = [c.c_void_p, P(DLTensor), c.c_int, c.c_float, c.c_float, P(Options), P(c.c_double), P(Stats)]
        lib.fnb_logits_cross_entropy.argtypes = [c.c_void_p, P(DLTensor), c.c_int, P(c.c_double)]
        lib.fnb_comm_unique_id.argtypes = [c.c_void_p]
        lib.fnb_comm_init.argtypes = [c.c_void_p, c.c_void_p, c.c_int, c.c_int]
        lib.fnb_comm_destroy.argtypes = [c.c_void_p]
        lib.fnb_comm_info.argtypes = [c.c_void_p, P(c.c_int), P(c.c_int), P(c.c_int)]
        lib.fnb_comm_last_error.argtypes = []
        lib.fnb_debug_chunk_plan.argtypes = [c.c_longlong, c.c_longlong, c.c_longlong, c.c_int, P(c.c_int), P(c.c_int)]
        lib.fnb_comm_shared_queue.argtypes = [c.c_void_p]
        lib.fnb_comm_last_error.restype = c.c_char_p
        lib.fnb_pair_histogram_sharded.argtypes = [c.c_void_p, P(DLTensor), P(DLTensor), P(c.c_double), c.c_int, P(Options),
                                                   P(DLTensor), P(Stats)]
        for name in EXPORTS:
            fn = getattr(lib, name)
            if fn.restype is c.c_int and name not in ('fnb_version',):
                fn.restype = c.c_int
        _lib = lib
        return lib


# ----------------------------------------------------------------------------------------
# DLPack ingestion

_pyapi = ctypes.pythonapi
_pyapi.PyCapsule_GetPointer.restype = ctypes.c_void_p
_pyapi.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
_pyapi.PyCapsule_IsValid.restype = ctypes.c_int
_pyapi.PyCapsule_IsValid.argtypes = [ctypes.py_object, ctypes.c_char_p]

_NP_CODES = {'i': 0, 'u': 1, 'f': 2}


class DLManagedTensor(ctypes.Structure):
    pass


DLManagedTensor._fields_ = [('dl_tensor', DLTensor), ('manager_ctx', ctypes.c_void_p),
                            ('deleter', ctypes.CFUNCTYPE(None, ctypes.POINTER(DLManagedTensor)))]

_pyapi.PyCapsule_SetName.restype = ctypes.c_int
_pyapi.PyCapsule_SetName.argtypes = [ctypes.py_object, ctypes.c_char_p]
_USED_NAME = b'used_dltensor'            # PyCapsule_SetName keeps the pointer, not a copy: module lifetime


def is_capsule(obj):
    return type(obj).__name__ == 'PyCapsule'


class DLPackTensor:
    """A tensor taken over from a raw ``"dltensor"`` PyCapsule -- what ``tf.experimental.dlpack.to_dlpack(t)`` and
    ``torch.utils.dlpack.to_dlpack(t)`` return -- per the DLPack protocol: the capsule is renamed ``"used_dltensor"`` (its
    producer will not free the tensor any more) and this object owns the ``DLManagedTensor``: its ``deleter`` runs when the
    object is released.  Exposes what the statistics classes need from an embedding matrix: ``shape``, ``dtype`` info,
    ``__dlpack_device__`` and the borrowed ``DLTensor*`` for the C ABI (the memory is used in place, never copied here)."""

    def __init__(self, capsule):
        if not is_capsule(capsule) or not _pyapi.PyCapsule_IsValid(capsule, b'dltensor'):
            raise ValueError('expected an un-consumed "dltensor" PyCapsule')
        addr = _pyapi.PyCapsule_GetPointer(capsule, b'dltensor')
        _pyapi.PyCapsule_SetName(capsule, _USED_NAME)
        self._managed = ctypes.cast(addr, ctypes.POINTER(DLManagedTensor))
        self._capsule = capsule                      # keeps the (now inert) capsule object alive with us
        t = self._managed.contents.dl_tensor
        self.ptr = ctypes.cast(addr, ctypes.POINTER(DLTensor))      # dl_tensor is the first member
        self.shape = tuple(int(t.shape[i]) for i in range(t.ndim))
        self.ndim = int(t.ndim)
        self.device = (int(t.device.device_type), int(t.device.device_id))
        self.is_cuda = self.device[0] == 2
        self.dtype_code, self.dtype_bits = int(t.dtype.code), int(t.dtype.bits)

    def __dlpack_device__(self):
        return self.device

    def __len__(self):
        return self.shape[0]

    def release(self):
        m, self._managed = self._managed, None
        if m is not None and m.contents.deleter:
            m.contents.deleter(m)

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def from_dlpack(obj):
    """``DLPackTensor`` for a raw capsule; anything else (NumPy, torch, objects with ``__dlpack__``) is returned unchanged."""
    return DLPackTensor(obj) if is_capsule(obj) else obj


class Borrowed:
    """A DLTensor borrowed from ``obj``; keeps whatever owns the memory alive."""

    def __init__(self, obj):
        self.keep = [obj]
        self.ptr = None
        capsule = None
        if is_capsule(obj):
            obj = DLPackTensor(obj)                  # consumed per protocol; lives as long as this borrow
            self.keep.append(obj)
        if isinstance(obj, DLPackTensor):
            self.ptr = obj.ptr
            return
        if hasattr(obj, '__dlpack__') and not (isinstance(obj, np.ndarray) and not obj.flags.writeable):
            try:
                capsule = obj.__dlpack__()
            except Exception:
                if not isinstance(obj, np.ndarray):
                    raise
        if capsule is not None:
            if not _pyapi.PyCapsule_IsValid(capsule, b'dltensor'):
                raise ValueError('object did not produce a "dltensor" capsule')
            addr = _pyapi.PyCapsule_GetPointer(capsule, b'dltensor')      # DLManagedTensor*, dl_tensor at offset 0
            self.keep.append(capsule)          # capsule stays un-consumed: its destructor calls the deleter
            self.ptr = ctypes.cast(addr, ctypes.POINTER(DLTensor))
        else:
            arr = np.asarray(obj)
            if not arr.flags.c_contiguous:
                raise ValueError('array must be C-contiguous')
            shape = (ctypes.c_int64 * max(arr.ndim, 1))(*arr.shape)
            t = DLTensor()
            t.data = arr.ctypes.data
            t.device = DLDevice(1, 0)
            t.ndim = arr.ndim
            t.dtype = DLDataType(_NP_CODES[arr.dtype.kind], arr.dtype.itemsize * 8, 1)
            t.shape = shape
            t.strides = None
            t.byte_offset = 0
            self.keep += [arr, shape, t]
            self.ptr = ctypes.pointer(t)


def _as_f32_matrix(x, name):
    """float32 C-contiguous [N, D]; NumPy inputs are converted if needed, GPU tensors must already comply."""
    if is_capsule(x):
        x = DLPackTensor(x)
    if isinstance(x, DLPackTensor):
        return x
    if isinstance(x, np.ndarray) or not hasattr(x, '__dlpack__'):
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2:
            raise ValueError('%s must be 2-D' % name)
    return x


def _as_labels(x):
    if is_capsule(x):
        x = DLPackTensor(x)
    if isinstance(x, DLPackTensor):
        return x
    if isinstance(x, np.ndarray) or not hasattr(x, '__dlpack__'):
        x = np.asarray(x)
        if x.dtype.kind not in 'iu' or x.dtype.itemsize not in (4, 8) or x.dtype.kind == 'u':
            x = x.astype(np.int64)
        x = np.ascontiguousarray(x)
    return x


class FnbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


_cuts_cache = {}


def numpy_cuts(thresholds, metric):
    """Cached front of ``_numpy_cuts`` (the bisection costs ~1 ms; every step of a validation loop asks for the same grid)."""
    thr = np.ascontiguousarray(np.atleast_1d(np.asarray(thresholds, dtype=np.float64)))
    key = (int(metric), thr.tobytes())
    hit = _cuts_cache.get(key)
    if hit is None:
        if len(_cuts_cache) > 64:
            _cuts_cache.clear()
        hit = _cuts_cache[key] = _numpy_cuts(thr, metric)
        hit.setflags(write=False)
    return hit


def _numpy_cuts(thresholds, metric):
    """Per threshold, the smallest float32 similarity s in [-1, 1] with
    ``dist(s) < threshold`` (float64 compare of the float32 distance, statistics.py:131), +inf if
    none -- evaluated with NumPy's own float32 arithmetic (``2 * (1 - s)`` / ``np.arccos``), so the
    kernel's similarity-domain bins reproduce NumPy's distance-domain comparison exactly."""
    thr = np.atleast_1d(np.asarray(thresholds, dtype=np.float64))

    def dist(s):
        s = np.clip(s, np.float32(-1), np.float32(1))
        return (2 * (1 - s)) if metric == 0 else np.arccos(s)

    def to_ord(f):
        u = np.asarray(f, dtype=np.float32).view(np.uint32).astype(np.int64)
        return np.where(u & 0x80000000, 0xFFFFFFFF - u, u | 0x80000000)

    def from_ord(o):
        o = np.asarray(o, dtype=np.int64)
        u = np.where(o & 0x80000000, o & 0x7FFFFFFF, 0xFFFFFFFF - o)
        return u.astype(np.uint32).view(np.float32)

    def pred(s):
        return dist(s.astype(np.float32)).astype(np.float64) < thr

    n = thr.size
    lo = np.full(n, to_ord(np.float32(-1)), dtype=np.int64)
    hi = np.full(n, to_ord(np.float32(1)), dtype=np.int64)
    none = ~pred(np.full(n, 1, dtype=np.float32))
    allp = pred(np.full(n, -1, dtype=np.float32))
    while np.any(hi - lo > 1):
        mid = lo + (hi - lo) // 2
        ok = pred(from_ord(mid))
        hi = np.where(ok, mid, hi)
        lo = np.where(ok, lo, mid)
    cuts = from_ord(hi).astype(np.float32)
    cuts[allp] = -1
    cuts[none] = np.inf
    return cuts


class Handle:
    """One library context (one GPU, one stream).  Not thread-safe."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = ctypes.c_void_p()
        rc = self.lib.fnb_create(int(device), ctypes.byref(h))
        if rc != FNB_OK:
            raise FnbError(rc, 'fnb_create(%d) failed: %s' % (device, self.lib.fnb_last_error(None).decode()))
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, 'h', None):
            self.lib.fnb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _raise(self, rc):
        raise FnbError(rc, self.lib.fnb_last_error(self.h).decode())

    def set_stream(self, cuda_stream=None):
        """Run on ``cuda_stream`` (an int handle, e.g. ``torch.cuda.current_stream().cuda_stream``; 0 is the legacy default
        stream); None = the handle's own stream."""
        arg = ctypes.c_void_p(-1) if cuda_stream is None else ctypes.c_void_p(int(cuda_stream))
        rc = self.lib.fnb_set_stream(self.h, arg)
        if rc != FNB_OK:
            self._raise(rc)
        self._stream = None if cuda_stream is None else int(cuda_stream)

    def _borrow(self, obj):
        """Borrow a tensor for one call.  A torch CUDA tensor is ready on torch's CURRENT stream (its producer kernels,
        an all-gather, a ``torch.zeros`` fill are queued there): the handle follows that stream so its work is ordered
        after them -- never a host synchronisation, never its own unordered stream."""
        if not isinstance(obj, np.ndarray) and getattr(obj, 'is_cuda', False):
            import torch
            dev = obj.device[1] if isinstance(obj, DLPackTensor) else obj.device
            s = int(torch.cuda.current_stream(dev).cuda_stream)
            if getattr(self, '_stream', None) != s:
                self.set_stream(s)
        return Borrowed(obj)

    def device_info(self):
        sm, ma, mi, mem = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_uint64()
        self.lib.fnb_device_info(self.h, ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi), ctypes.byref(mem))
        return {'sm_count': sm.value, 'cc': (ma.value, mi.value), 'total_mem': mem.value}

    def options(self, mode='fp16x3', metric=0, atol=1.e-5, eps=1.e-5, rank=0, world=1, cta_group=0, region_rows=0,
                cuts=None, max_ctas=0, force_checked=False, cluster_pairs=0, normalize=0, theta=0.0, raw_distance=False,
                shard=None, subset_rows=0, panel_window=None, strict_tiles=None, bias_correction=None, streamed=None, tile_queue=None):
        o = Options()
        self.lib.fnb_default_options(ctypes.byref(o))
        o.mode = MODES[mode] if isinstance(mode, str) else int(mode)
        o.metric = int(metric)
        o.atol = float(atol)
        o.eps = float(eps)
        o.rank, o.world = int(rank), int(world)
        o.cta_group = int(cta_group)
        o.region_rows = int(region_rows)
        o.max_ctas = int(max_ctas)
        o.force_checked = 1 if force_checked else 0
        o.debug = int(os.environ.get('FNB_DEBUG', '0'))     # profiling knob (see fnb_options.debug)
        o.cluster_pairs = int(cluster_pairs)
        o.normalize = int(normalize)
        o.theta = float(theta)
        o.raw_distance = 1 if raw_distance else 0
        o.subset_rows = int(subset_rows)
        # None: FNB_PANEL_WINDOW from the environment (probe scripts), else 0 = auto (see fnb_options.panel_window)
        o.panel_window = int(os.environ.get('FNB_PANEL_WINDOW', '0')) if panel_window is None else int(panel_window)
        # measurement knobs (None: environment, else 0 = on, -1 = off)
        o.strict_tiles = int(os.environ.get('FNB_STRICT_TILES', '0')) if strict_tiles is None else int(strict_tiles)
        o.bias_correction = int(os.environ.get('FNB_BIAS_CORRECTION', '0')) if bias_correction is None else int(bias_correction)
        # chunked upload under the launches for host rows in class order: 0 = auto, 1 = always, -1 = off (fnb_options.streamed)
        o.streamed = int(os.environ.get('FNB_STREAMED', '0')) if streamed is None else int(streamed)
        # tile queue of the histogram launches: 0 = on (+ the second launch on the free SMs), 2 = queue only, -1 = static schedule
        o.tile_queue = int(os.environ.get('FNB_TILE_QUEUE', '0')) if tile_queue is None else int(tile_queue)
        keep = None
        if cuts is not None:
            keep = np.ascontiguousarray(cuts, dtype=np.float32)
            o.cuts = keep.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        if shard is not None:
            # (mod, lo, width): a contiguous residue range; (mod, [residues]): an explicit ascending list
            if len(shard) == 2:
                slots = np.ascontiguousarray(shard[1], dtype=np.int32)
                o.shard_mod, o.shard_lo, o.shard_width = int(shard[0]), 0, int(slots.size)
                o.shard_slots = slots.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
                keep = (keep, slots)
            else:
                o.shard_mod, o.shard_lo, o.shard_width = (int(v) for v in shard)
        return o, keep

    # ---- pairwise_similarities (statistics.py:22-57)
    def pairwise(self, xa, xb=None, metric=0, atol=1.e-5, mode='fp16x3', cta_group=0, out=None, normalize=0, theta=0.0,
                 raw_distance=False, bias_correction=None):
        xa = _as_f32_matrix(xa, 'xa')
        na = xa.shape[0]
        if xb is not None:
            xb = _as_f32_matrix(xb, 'xb')
            shape = (na, xb.shape[0])
        else:
            shape = (na * (na - 1) // 2,)
        if out is None:
            out = np.empty(shape, dtype=np.float32)
        o, keep = self.options(mode=mode, metric=metric, atol=atol, cta_group=cta_group, normalize=normalize, theta=theta,
                               raw_distance=raw_distance, bias_correction=bias_correction)
        rng = (ctypes.c_float * 2)()
        ba, bo = self._borrow(xa), self._borrow(out)
        bb = self._borrow(xb) if xb is not None else None
        rc = self.lib.fnb_pairwise(self.h, ba.ptr, bb.ptr if bb else None, ctypes.byref(o), bo.ptr, rng)
        self.last_range = (rng[0], rng[1])
        if rc != FNB_OK:
            self._raise(rc)
        return out

    # ---- whole-set verification histogram
    def pair_histogram_bins(self, embeddings, labels, thresholds, metric=0, atol=1.e-5, eps=1.e-5, mode='fp16x3',
                            rank=0, world=1, cta_group=0, region_rows=0, bins_out=None, max_ctas=0, cuts='numpy',
                            force_checked=False, cluster_pairs=0, normalize=0, shard=None, panel_window=None,
                            strict_tiles=None, bias_correction=None, streamed=None, tile_queue=None):
        embeddings = _as_f32_matrix(embeddings, 'embeddings')
        labels = _as_labels(labels)
        thr = np.ascontiguousarray(np.atleast_1d(thresholds), dtype=np.float64)
        if isinstance(cuts, str) and cuts == 'numpy':
            cuts = numpy_cuts(thr, metric)
        o, keep = self.options(mode=mode, metric=metric, atol=atol, eps=eps, rank=rank, world=world, cta_group=cta_group,
                               region_rows=region_rows, cuts=cuts, max_ctas=max_ctas, force_checked=force_checked,
                               cluster_pairs=cluster_pairs, normalize=normalize, shard=shard, panel_window=panel_window,
                               strict_tiles=strict_tiles, bias_correction=bias_correction, streamed=streamed, tile_queue=tile_queue)
        if bins_out is None:
            bins_out = np.zeros((2, thr.size + 1), dtype=np.uint64)
        st = Stats()
        be, bl, bb = self._borrow(embeddings), self._borrow(labels), self._borrow(bins_out)
        rc = self.lib.fnb_pair_histogram_bins(self.h, be.ptr, bl.ptr, thr.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                              thr.size, ctypes.byref(o), bb.ptr, ctypes.byref(st))
        if rc != FNB_OK:
            self._raise(rc)
        return bins_out, st.as_dict()

    # ---- multi-GPU inside the library (NCCL behind the C ABI)
    def comm_unique_id(self):
        """128 bytes (an ncclUniqueId) made on rank 0; ship them to the other ranks and call ``comm_init`` everywhere."""
        buf = ctypes.create_string_buffer(128)
        rc = self.lib.fnb_comm_unique_id(buf)
        if rc != FNB_OK:
            raise FnbError(rc, self.lib.fnb_comm_last_error().decode())
        return buf.raw

    def comm_init(self, unique_id, rank, world):
        rc = self.lib.fnb_comm_init(self.h, ctypes.c_char_p(bytes(unique_id)), int(rank), int(world))
        if rc != FNB_OK:
            self._raise(rc)
        self.comm = (int(rank), int(world))

    def comm_init_from_torch(self, group=None):
        """Communicator over the ranks of a ``torch.distributed`` group: rank 0's id travels through the group's own
        broadcast (any backend); the communicator itself is NCCL's, inside the library."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [self.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self.comm_init(box[0], rank, world)

    def comm_destroy(self):
        if getattr(self, 'comm', None) is not None:
            self.lib.fnb_comm_destroy(self.h)
            self.comm = None

    def comm_info(self):
        r, w, v = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self.lib.fnb_comm_info(self.h, ctypes.byref(r), ctypes.byref(w), ctypes.byref(v))
        return {'rank': r.value, 'world': w.value, 'nccl_version': v.value, 'shared_queue': bool(self.lib.fnb_comm_shared_queue(self.h))}

    def pair_histogram_sharded(self, emb_shard, labels_shard, thresholds, metric=0, atol=1.e-5, eps=1.e-5, mode='auto',
                               cta_group=0, region_rows=0, bins_out=None, cuts='numpy', cluster_pairs=0, normalize=0, shard=None,
                               panel_window=None, strict_tiles=None, bias_correction=None, streamed=None):
        """Collective over the handle's communicator: every rank passes its rows, every rank gets the bins of the whole set
        (``fnb_pair_histogram_sharded``)."""
        emb_shard = _as_f32_matrix(emb_shard, 'embeddings')
        labels_shard = _as_labels(labels_shard)
        thr = np.ascontiguousarray(np.atleast_1d(thresholds), dtype=np.float64)
        if isinstance(cuts, str) and cuts == 'numpy':
            cuts = numpy_cuts(thr, metric)
        o, keep = self.options(mode=mode, metric=metric, atol=atol, eps=eps, cta_group=cta_group, region_rows=region_rows, cuts=cuts,
                               cluster_pairs=cluster_pairs, normalize=normalize, shard=shard, panel_window=panel_window,
                               strict_tiles=strict_tiles, bias_correction=bias_correction, streamed=streamed)
        if bins_out is None:
            bins_out = np.zeros((2, thr.size + 1), dtype=np.uint64)
        st = Stats()
        be, bl, bb = self._borrow(emb_shard), self._borrow(labels_shard), self._borrow(bins_out)
        rc = self.lib.fnb_pair_histogram_sharded(self.h, be.ptr, bl.ptr, thr.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                                 thr.size, ctypes.byref(o), bb.ptr, ctypes.byref(st))
        if rc != FNB_OK:
            self._raise(rc)
        return bins_out, st.as_dict()

    def counts_from_bins(self, bins, thresholds, metric=0, eps=1.e-5, cuts='numpy'):
        thr = np.ascontiguousarray(np.atleast_1d(thresholds), dtype=np.float64)
        if isinstance(cuts, str) and cuts == 'numpy':
            cuts = numpy_cuts(thr, metric)
        o, keep = self.options(metric=metric, eps=eps, cuts=cuts)
        bins = np.ascontiguousarray(bins, dtype=np.uint64)
        same = np.zeros(thr.size, dtype=np.uint64)
        diff = np.zeros(thr.size, dtype=np.uint64)
        ns, nd = ctypes.c_uint64(), ctypes.c_uint64()
        u64p = ctypes.POINTER(ctypes.c_uint64)
        rc = self.lib.fnb_counts_from_bins(thr.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), thr.size, ctypes.byref(o),
                                           bins.ctypes.data_as(u64p), same.ctypes.data_as(u64p), diff.ctypes.data_as(u64p),
                                           ctypes.byref(ns), ctypes.byref(nd))
        if rc != FNB_OK:
            raise FnbError(rc, 'fnb_counts_from_bins failed (%d)' % rc)
        return {'same': same.astype(np.int64), 'diff': diff.astype(np.int64), 'n_same': ns.value, 'n_diff': nd.value}

    def pair_histogram(self, embeddings, labels, thresholds, metric=0, **kw):
        eps = kw.get('eps', 1.e-5)
        bins, st = self.pair_histogram_bins(embeddings, labels, thresholds, metric=metric, **kw)
        out = self.counts_from_bins(bins, thresholds, metric=metric, eps=eps)
        out['stats'] = st
        out['bins'] = bins
        return out

    # ---- keyed histogram over rectangles
    def region_histogram_bins(self, embeddings, perm, cls, regions, nkeys, thresholds, metric=0, atol=1.e-5, eps=1.e-5,
                              mode='fp16x3', cta_group=0, cuts='numpy', rank=0, world=1, cluster_pairs=0, normalize=0, theta=0.0,
                              raw_distance=False, subset=False):
        """``subset=True``: ``perm`` / ``cls`` describe ``len(perm)`` rows of ``embeddings`` (``perm`` indexes the full array),
        e.g. one fold of a validation whose embeddings stay resident on the GPU."""
        embeddings = _as_f32_matrix(embeddings, 'embeddings')
        perm = np.ascontiguousarray(perm, dtype=np.int64)
        cls = np.ascontiguousarray(cls, dtype=np.int32)
        regions = np.ascontiguousarray(regions, dtype=REGION_DTYPE)
        thr = np.ascontiguousarray(np.atleast_1d(thresholds), dtype=np.float64)
        if isinstance(cuts, str) and cuts == 'numpy':
            cuts = numpy_cuts(thr, metric)
        o, keep = self.options(mode=mode, metric=metric, atol=atol, eps=eps, cta_group=cta_group, cuts=cuts, rank=rank, world=world,
                               cluster_pairs=cluster_pairs, normalize=normalize, theta=theta, raw_distance=raw_distance,
                               subset_rows=perm.size if subset else 0)
        if cls.size != perm.size or (not subset and perm.size != int(embeddings.shape[0])):
            raise ValueError('perm / cls must have one entry per row')
        bins = np.zeros((int(nkeys), 2, thr.size + 1), dtype=np.uint64)
        st = Stats()
        be = self._borrow(embeddings)
        rc = self.lib.fnb_region_histogram_bins(
            self.h, be.ptr, perm.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), cls.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
            regions.ctypes.data_as(ctypes.POINTER(Region)), regions.size, int(nkeys),
            thr.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), thr.size, ctypes.byref(o),
            bins.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), ctypes.byref(st))
        if rc != FNB_OK:
            self._raise(rc)
        return bins, st.as_dict()


    # ---- class-balanced confidence matrix + threshold selection from the bins left on the device
    def confidence_from_last_bins(self, nkeys, w_same, w_diff, thresholds, metric=0, far_target=0.0, cuts='numpy', eps=1.e-5):
        """tp, tn, fp, fn (float64 [T]), index of the first accuracy maximum and the FAR threshold
        (statistics.py:130-138,296,299-302) from the keyed bins of the last ``region_histogram_bins`` call."""
        thr = np.ascontiguousarray(np.atleast_1d(thresholds), dtype=np.float64)
        if isinstance(cuts, str) and cuts == 'numpy':
            cuts = numpy_cuts(thr, metric)
        o, keep = self.options(metric=metric, eps=eps, cuts=cuts)
        ws = np.ascontiguousarray(w_same, dtype=np.float64).reshape(-1)
        wd = np.ascontiguousarray(w_diff, dtype=np.float64).reshape(-1)
        if ws.size != nkeys or wd.size != nkeys:
            raise ValueError('w_same / w_diff must have nkeys entries')
        out = np.zeros((4, thr.size), dtype=np.float64)
        amax = ctypes.c_int32(-1)
        far = ctypes.c_double(0.0)
        dp = ctypes.POINTER(ctypes.c_double)
        rc = self.lib.fnb_confidence_from_last_bins(
            self.h, int(nkeys), ws.ctypes.data_as(dp), wd.ctypes.data_as(dp), thr.ctypes.data_as(dp), thr.size,
            ctypes.byref(o), float(far_target), out[0].ctypes.data_as(dp), out[1].ctypes.data_as(dp),
            out[2].ctypes.data_as(dp), out[3].ctypes.data_as(dp), ctypes.byref(amax), ctypes.byref(far))
        if rc != FNB_OK:
            self._raise(rc)
        return {'tp': out[0], 'tn': out[1], 'fp': out[2], 'fn': out[3], 'argmax_accuracy': int(amax.value),
                'far_threshold': float(far.value)}

    # ---- pair-classifier cross entropy over one P x K batch (train_classifier.py:60-84)
    def pair_cross_entropy(self, batch, examples_per_class, alpha, threshold, normalize=0, theta=0.0):
        batch = _as_f32_matrix(batch, 'batch')
        o, keep = self.options(normalize=normalize, theta=theta, raw_distance=True)
        out = (ctypes.c_double * 5)()
        st = Stats()
        bb = self._borrow(batch)
        rc = self.lib.fnb_pair_cross_entropy(self.h, bb.ptr, int(examples_per_class), ctypes.c_float(alpha), ctypes.c_float(threshold),
                                             ctypes.byref(o), out, ctypes.byref(st))
        if rc != FNB_OK:
            self._raise(rc)
        return {'loss': out[0], 'dalpha': out[1], 'dthreshold': out[2], 'dtheta': out[3], 'pos_weight': out[4],
                'stats': st.as_dict()}

    def logits_cross_entropy(self, logits, examples_per_class):
        logits = _as_f32_matrix(logits, 'logits')
        loss = ctypes.c_double(0.0)
        bl = self._borrow(logits)
        rc = self.lib.fnb_logits_cross_entropy(self.h, bl.ptr, int(examples_per_class), ctypes.byref(loss))
        if rc != FNB_OK:
            self._raise(rc)
        return float(loss.value)

    # ---- triplet mining (semantics: oracle/mining_oracle.py; not in the reference fork)
    _kmax_cache = {}

    @classmethod
    def _kmax_of(cls, labels, nbatches=1):
        """largest class size - 1 over the batches (host labels; cached on the label bytes: a training loop re-uses its P x K labels)"""
        lab = np.ascontiguousarray(labels)
        key = (lab.shape, lab.dtype.str, int(nbatches), hash(lab.tobytes()))
        hit = cls._kmax_cache.get(key)
        if hit is None:
            if len(cls._kmax_cache) > 256:
                cls._kmax_cache.clear()
            hit = 0
            for part in np.split(lab, int(nbatches)) if lab.size else []:
                hit = max(hit, int(np.unique(part, return_counts=True)[1].max()) - 1)
            cls._kmax_cache[key] = hit
        return hit

    def mine(self, embeddings, labels, alpha=0.2, kmax=None, mode='fp16x3', atol=1.e-5):
        """One batch, host (NumPy) results: the synchronous form (``fnb_mine``)."""
        embeddings = _as_f32_matrix(embeddings, 'embeddings')
        labels = _as_labels(labels)
        b = int(embeddings.shape[0])
        if kmax is None:
            kmax = self._kmax_of(labels if isinstance(labels, np.ndarray) else np.asarray(labels.cpu())) if b else 0
        kmax = int(kmax)
        hp = np.full(b, -1, dtype=np.int32)
        hn = np.full(b, -1, dtype=np.int32)
        pi = np.full((b, kmax), -1, dtype=np.int32)
        sh = np.full((b, kmax), -1, dtype=np.int32)
        el = np.zeros((b, kmax), dtype=np.int32)
        o, keep = self.options(mode=mode, metric=0, atol=atol)
        st = Stats()
        be, bl = self._borrow(embeddings), self._borrow(labels)
        ip = ctypes.POINTER(ctypes.c_int32)
        rc = self.lib.fnb_mine(self.h, be.ptr, bl.ptr, float(alpha), ctypes.byref(o), hp.ctypes.data_as(ip), hn.ctypes.data_as(ip),
                               kmax, pi.ctypes.data_as(ip), sh.ctypes.data_as(ip), el.ctypes.data_as(ip), ctypes.byref(st))
        if rc != FNB_OK:
            self._raise(rc)
        return {'hardest_pos': hp, 'hardest_neg': hn, 'pos_index': pi, 'semi_hard': sh, 'eligible': el, 'stats': st.as_dict()}

    def mine_batched(self, embeddings, labels, nbatches=1, alpha=0.2, kmax=None, mode='fp16x3', atol=1.e-5, out=None):
        """``nbatches`` batches of equal size packed as ``[S * B, D]`` / ``[S * B]`` in ONE call (``fnb_mine_batched``); all indices
        are local to their batch.  torch CUDA inputs give torch CUDA int32 outputs WITHOUT a host synchronisation (pass ``kmax``;
        ``out`` = the dict of a previous call re-uses its tensors; data-dependent errors surface in ``mine_check``); NumPy inputs
        give NumPy outputs (synchronous).  ``kmax=0``: hardest positive / negative only -- fully fused in the Gram epilogue, no
        distance strip."""
        embeddings = _as_f32_matrix(embeddings, 'embeddings')
        labels = _as_labels(labels)
        rows = int(embeddings.shape[0])
        on_gpu = not isinstance(embeddings, np.ndarray) and getattr(embeddings, 'is_cuda', False)
        if kmax is None:
            kmax = self._kmax_of(labels if isinstance(labels, np.ndarray) else np.asarray(labels.cpu()), nbatches) if rows else 0
        kmax = int(kmax)
        names = ('hardest_pos', 'hardest_neg', 'pos_index', 'semi_hard', 'eligible', 'status')
        shapes = ((rows,), (rows,), (rows, kmax), (rows, kmax), (rows, kmax), (4,))
        if out is None:
            out = {}
        for name, shape in zip(names, shapes):
            cur = out.get(name)
            if cur is not None and tuple(cur.shape) == shape:
                continue
            if on_gpu:
                import torch
                out[name] = torch.empty(shape, dtype=torch.int32, device=embeddings.device)
            else:
                out[name] = np.empty(shape, dtype=np.int32)
        o, keep = self.options(mode=mode, metric=0, atol=atol)
        st = Stats()
        be, bl = self._borrow(embeddings), self._borrow(labels)
        bo = [self._borrow(out[name]) for name in names]
        none = ctypes.POINTER(DLTensor)()
        ptrs = [b.ptr for b in bo]
        if kmax == 0:
            ptrs[2] = ptrs[3] = ptrs[4] = none
        rc = self.lib.fnb_mine_batched(self.h, be.ptr, bl.ptr, int(nbatches), float(alpha), ctypes.byref(o), kmax,
                                       ptrs[0], ptrs[1], ptrs[2], ptrs[3], ptrs[4], ptrs[5], ctypes.byref(st))
        if rc != FNB_OK:
            self._raise(rc)
        out['stats'] = st.as_dict()
        return out

    def false_pairs(self, embeddings, labels, threshold, metric=0, atol=1.e-5, capacity=1 << 20, raw_distance=False):
        """Every same-identity pair with ``d > threshold`` and every different-identity pair with ``d < threshold``
        (``fnb_false_pairs``): ``(rows, cols, dist, same)`` with original row indices; grows the buffer and repeats when the list
        did not fit."""
        embeddings = _as_f32_matrix(embeddings, 'embeddings')
        labels = _as_labels(labels)
        o, keep = self.options(metric=metric, atol=atol, raw_distance=raw_distance)
        be, bl = self._borrow(embeddings), self._borrow(labels)
        while True:
            rows = np.empty(capacity, dtype=np.int32)
            cols = np.empty(capacity, dtype=np.int32)
            dist = np.empty(capacity, dtype=np.float32)
            count = ctypes.c_uint64(0)
            st = Stats()
            rc = self.lib.fnb_false_pairs(self.h, be.ptr, bl.ptr, float(threshold), ctypes.byref(o), int(capacity),
                                          rows.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), cols.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                          dist.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ctypes.byref(count), ctypes.byref(st))
            if rc != FNB_OK:
                self._raise(rc)
            if count.value <= capacity:
                break
            capacity = int(count.value) + 1024
        k = int(count.value)
        return rows[:k], cols[:k], dist[:k], st.as_dict()

    def mine_check(self):
        """Synchronise and raise what the last ``mine_batched`` call on device tensors could not report; returns
        ``{'kmax_needed', 'smin', 'smax', 'kernel_ms', 'prepare_ms'}``."""
        status = (ctypes.c_int32 * 4)()
        st = Stats()
        rc = self.lib.fnb_mine_check(self.h, status, ctypes.byref(st))
        if rc != FNB_OK:
            self._raise(rc)
        return {'kmax_needed': int(status[0]), 'smin': st.smin, 'smax': st.smax, 'kernel_ms': st.kernel_ms, 'prepare_ms': st.prepare_ms}

    def mine_select_kth(self, anchors, positives, kth, alpha=0.2, out=None):
        """The ``kth[i]``-th (0-based, ascending index) margin-eligible negative of ``(anchors[i], positives[i])`` over the distance
        strips of the last mining call (``fnb_mine_select_kth``); -1 if there are fewer.  int32 arrays (NumPy or torch CUDA)."""
        def as_i32(x):
            if isinstance(x, np.ndarray) or not hasattr(x, '__dlpack__'):
                return np.ascontiguousarray(x, dtype=np.int32)
            return x
        anchors, positives, kth = as_i32(anchors), as_i32(positives), as_i32(kth)
        if out is None:
            if isinstance(anchors, np.ndarray):
                out = np.empty(anchors.shape, dtype=np.int32)
            else:
                import torch
                out = torch.empty(anchors.shape, dtype=torch.int32, device=anchors.device)
        ba, bp, bk, bo = self._borrow(anchors), self._borrow(positives), self._borrow(kth), self._borrow(out)
        rc = self.lib.fnb_mine_select_kth(self.h, ba.ptr, bp.ptr, bk.ptr, float(alpha), bo.ptr)
        if rc != FNB_OK:
            self._raise(rc)
        return out


_default_handles = {}



def default_handle(device=0):
    """Process-wide handle per device (created on first use)."""
    h = _default_handles.get(device)
    if h is None or h.h is None:
        h = Handle(device)
        _default_handles[device] = h
    return h
