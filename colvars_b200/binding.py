"""ctypes binding of colvars_b200/libb200cv.so (the C ABI in include/b200cv.h).

Python is plumbing here: it loads the C-ABI library, hands raw device pointers of torch
tensors to it and mirrors the reference's component interface
(`colvar::cvc`: init / read_data / calc_value / calc_gradients / apply_force,
src/colvarcomp.h:108-229) closely enough that the parity tests read like the reference's own
functional tests.  There is no Python or CPU compute path: if the CUDA library is missing
or no GPU is present, construction raises.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
# B200CV_LIB selects another build of the same library (kernel-geometry experiments)
LIB_PATH = Path(os.environ.get("B200CV_LIB", _HERE / "libb200cv.so"))

OK = 0
ERROR = 1
NOT_IMPLEMENTED = 1 << 1
INPUT_ERROR = 1 << 2
BUG_ERROR = 1 << 3
FILE_ERROR = 1 << 4
MEMORY_ERROR = 1 << 5

COORDNUM, SELFCOORDNUM, GROUPCOORD, DISTANCEINV, DISTANCEPAIRS = 0, 1, 2, 3, 4
KIND_BY_NAME = {"coordNum": COORDNUM, "selfCoordNum": SELFCOORDNUM, "groupCoord": GROUPCOORD,
                "distanceInv": DISTANCEINV, "distancePairs": DISTANCEPAIRS}

EF_NULL = 0
EF_GRADIENTS = 1
EF_USE_PAIRLIST = 1 << 9
EF_REBUILD_PAIRLIST = 1 << 10


class Config(C.Structure):
    _fields_ = [
        ("kind", C.c_int),
        ("n1", C.c_int),
        ("n2", C.c_int),
        ("cutoff3", C.c_double * 3),
        ("exp_numer", C.c_int),
        ("exp_denom", C.c_int),
        ("group1_center_only", C.c_int),
        ("group2_center_only", C.c_int),
        ("tolerance", C.c_double),
        ("pairlist_freq", C.c_int),
        ("device", C.c_int),
    ]


class B200cvError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200cv error {code}: {msg}")
        self.code = code


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

# every symbol include/b200cv.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "b200cv_config_defaults", "b200cv_create", "b200cv_destroy", "b200cv_last_error",
    "b200cv_set_group", "b200cv_set_fitting", "b200cv_set_boundaries", "b200cv_read_data",
    "b200cv_calc_value", "b200cv_step", "b200cv_calc_fit_gradients", "b200cv_value", "b200cv_apply_force",
    "b200cv_eval", "b200cv_eval_async", "b200cv_accumulate_gradients",
    "b200cv_pairlist_reserve", "b200cv_pairlist_set_rebuild", "b200cv_async_set_skip", "b200cv_eval_host", "b200cv_eval_pairs_host", "b200cv_distance_pairs_host", "b200cv_distance_pairs_force_host", "b200cv_calc_host", "b200cv_apply_force_host", "b200cv_get_positions",
    "b200cv_get_gradients", "b200cv_get_fit_gradients", "b200cv_get_centers",
    "b200cv_get_rotation", "b200cv_get_tolerance_l2_max", "b200cv_pairlist_count",
    "b200cv_pairlist_export", "b200cv_pairlist_stats", "b200cv_last_stats", "b200cv_kernel_launches", "b200cv_current_device", "b200cv_comm_unique_id",
    "b200cv_comm_init", "b200cv_set_shard", "b200cv_plan_items", "b200cv_fp64_peak", "b200cv_selftest_rcp",
]


def load():
    """Load libb200cv.so.  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(str(LIB_PATH), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else 0)
    vp, sz, ll = C.c_void_p, C.c_size_t, C.c_longlong
    H = C.c_void_p
    L.b200cv_config_defaults.argtypes = [C.POINTER(Config)]
    L.b200cv_config_defaults.restype = None
    L.b200cv_create.argtypes = [C.POINTER(Config), C.POINTER(H)]
    L.b200cv_destroy.argtypes = [H]
    L.b200cv_last_error.argtypes = [H]
    L.b200cv_last_error.restype = C.c_char_p
    L.b200cv_set_group.argtypes = [H, C.c_int, _ip, _dp, C.c_int]
    L.b200cv_set_fitting.argtypes = [H, C.c_int, C.c_int, C.c_int, _ip, C.c_int, _dp, C.c_int]
    L.b200cv_set_boundaries.argtypes = [H, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]
    L.b200cv_read_data.argtypes = [H, vp, sz, vp]
    L.b200cv_calc_value.argtypes = [H, ll, C.c_int, vp]
    L.b200cv_step.argtypes = [H, vp, sz, ll, C.c_int, vp]
    L.b200cv_calc_fit_gradients.argtypes = [H, vp]
    L.b200cv_value.argtypes = [H, _dp]
    L.b200cv_apply_force.argtypes = [H, C.c_double, vp, sz, vp]
    L.b200cv_eval.argtypes = [H, vp, vp, C.c_int, vp, vp, vp]
    L.b200cv_eval_async.argtypes = [H, vp, vp, C.c_int, vp, vp, vp]
    L.b200cv_accumulate_gradients.argtypes = [H, vp, vp, vp]
    L.b200cv_pairlist_reserve.argtypes = [H, vp, vp, C.c_double, vp]
    L.b200cv_pairlist_set_rebuild.argtypes = [H, C.c_int]
    L.b200cv_async_set_skip.argtypes = [H, C.c_int]
    L.b200cv_eval_host.argtypes = [H, _dp, _dp, C.c_int, _dp, _dp, _dp]
    L.b200cv_eval_pairs_host.argtypes = [H, _dp, _dp, C.c_int, _dp, _dp, _dp]
    L.b200cv_distance_pairs_host.argtypes = [H, _dp, _dp, _dp]
    L.b200cv_distance_pairs_force_host.argtypes = [H, _dp, _dp, _dp]
    L.b200cv_calc_host.argtypes = [H, vp, sz, ll, C.c_int, _dp]
    L.b200cv_apply_force_host.argtypes = [H, C.c_double, vp, sz]
    L.b200cv_get_positions.argtypes = [H, C.c_int, _dp]
    L.b200cv_get_gradients.argtypes = [H, C.c_int, _dp]
    L.b200cv_get_fit_gradients.argtypes = [H, C.c_int, _dp, C.c_int]
    L.b200cv_get_centers.argtypes = [H, C.c_int, _dp, _dp]
    L.b200cv_get_rotation.argtypes = [H, C.c_int, _dp]
    L.b200cv_get_tolerance_l2_max.argtypes = [H, _dp]
    L.b200cv_pairlist_count.argtypes = [H, C.POINTER(sz)]
    L.b200cv_pairlist_export.argtypes = [H, C.POINTER(C.c_uint64), sz]
    L.b200cv_pairlist_stats.argtypes = [H, _ip, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz),
                                        C.POINTER(sz)]
    L.b200cv_last_stats.argtypes = [H, _ip, C.POINTER(sz)]
    L.b200cv_current_device.argtypes = [_ip]
    L.b200cv_kernel_launches.argtypes = []
    L.b200cv_kernel_launches.restype = C.c_ulonglong
    L.b200cv_comm_unique_id.argtypes = [C.c_char_p]
    L.b200cv_comm_init.argtypes = [H, C.c_char_p, C.c_int, C.c_int]
    L.b200cv_set_shard.argtypes = [H, C.c_int, C.c_int]
    L.b200cv_plan_items.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, sz,
                                    C.POINTER(sz), _ip]
    L.b200cv_fp64_peak.argtypes = [C.c_int, C.c_double, _dp, _dp]
    L.b200cv_selftest_rcp.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, _dp]
    for name in SYMBOLS:
        fn = getattr(L, name)  # raises AttributeError when a declared symbol is missing
        if name not in ("b200cv_last_error", "b200cv_config_defaults"):
            fn.restype = C.c_int
    _lib = L
    return L


def _as_dp(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _stream_ptr(stream):
    """cudaStream_t of a torch stream (None -> torch's current stream)."""
    import torch
    if stream is None:
        stream = torch.cuda.current_stream()
    return C.c_void_p(stream.cuda_stream)


class CoordNum:
    """One coordination-number component behind the C ABI; method names follow colvar::cvc.

    group1 / group2: proxy slots of the atoms (0-based), mass1 / mass2 optional.
    """

    def __init__(self, function_type="coordNum", group1=None, group2=None, *, cutoff=None,
                 cutoff3=None, exp_numer=6, exp_denom=12, group1_center_only=False,
                 group2_center_only=False, tolerance=0.0, pairlist_freq=100, mass1=None,
                 mass2=None, device=0, exponent=None):
        L = load()
        self._L = L
        self._h = C.c_void_p()
        cfg = Config()
        L.b200cv_config_defaults(C.byref(cfg))
        cfg.kind = KIND_BY_NAME[function_type]
        self.function_type = function_type
        group1 = np.ascontiguousarray(group1, dtype=np.int32)
        self.n1 = len(group1)
        cfg.n1 = self.n1
        if function_type != "selfCoordNum":
            group2 = np.ascontiguousarray(group2, dtype=np.int32)
            self.n2 = len(group2)
            cfg.n2 = self.n2
        else:
            self.n2 = 0
        if cutoff is not None and cutoff3 is not None:
            raise B200cvError(INPUT_ERROR,
                              'Error: cannot specify "cutoff" and "cutoff3" at the same time.')
        if cutoff is not None:
            cutoff3 = (cutoff, cutoff, cutoff)
        if cutoff3 is not None:
            for d in range(3):
                cfg.cutoff3[d] = float(cutoff3[d])
        cfg.exp_numer, cfg.exp_denom = int(exp_numer), int(exp_denom)
        if function_type == "distanceInv":
            # keyword "exponent" (src/colvarcomp_distances.cpp:383), default 6
            cfg.exp_numer = 6 if exponent is None else int(exponent)
        cfg.group1_center_only = int(bool(group1_center_only))
        cfg.group2_center_only = int(bool(group2_center_only))
        cfg.tolerance = float(tolerance)
        cfg.pairlist_freq = int(pairlist_freq)
        cfg.device = int(device)
        self.device = int(device)
        rc = L.b200cv_create(C.byref(cfg), C.byref(self._h))
        if rc != OK:
            raise B200cvError(rc, L.b200cv_last_error(None).decode())
        self.cfg = cfg
        m1 = None if mass1 is None else np.ascontiguousarray(mass1, dtype=np.float64)
        self._check(L.b200cv_set_group(self._h, 1, group1.ctypes.data_as(_ip), _as_dp(m1),
                                       self.n1))
        if function_type != "selfCoordNum":
            m2 = None if mass2 is None else np.ascontiguousarray(mass2, dtype=np.float64)
            self._check(L.b200cv_set_group(self._h, 2, group2.ctypes.data_as(_ip), _as_dp(m2),
                                           self.n2))

    # -- plumbing ------------------------------------------------------------------------
    def _check(self, rc):
        if rc != OK:
            raise B200cvError(rc, self._L.b200cv_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.b200cv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration --------------------------------------------------------------------
    def set_boundaries(self, periodic, A, B, Cc):
        A, B, Cc = (np.ascontiguousarray(v, dtype=np.float64) for v in (A, B, Cc))
        self._check(self._L.b200cv_set_boundaries(self._h, int(periodic[0]), int(periodic[1]),
                                                  int(periodic[2]), _as_dp(A), _as_dp(B),
                                                  _as_dp(Cc)))

    def set_fitting(self, which, *, center=True, rotate=True, fitting_group=None, ref_pos=None,
                    enable_fit_gradients=True):
        """ref_pos: (nref, 3) reference positions (refPositions)."""
        fg = None if fitting_group is None else np.ascontiguousarray(fitting_group, np.int32)
        nfit = 0 if fg is None else len(fg)
        rp = np.ascontiguousarray(np.asarray(ref_pos, dtype=np.float64).T).reshape(-1)
        self._check(self._L.b200cv_set_fitting(
            self._h, which, int(center), int(rotate),
            None if fg is None else fg.ctypes.data_as(_ip), nfit, _as_dp(rp),
            int(enable_fit_gradients)))

    def comm_init(self, unique_id, rank, world):
        self._check(self._L.b200cv_comm_init(self._h, unique_id, rank, world))

    def set_shard(self, rank, world):
        self._check(self._L.b200cv_set_shard(self._h, rank, world))

    # -- device-resident step (torch tensors, colvarproxy_gpu SoA layout) ----------------
    def read_data(self, proxy_pos, stream=None):
        """proxy_pos: torch float64 CUDA tensor (3, P) = SoA planes with stride P."""
        assert proxy_pos.is_cuda and proxy_pos.dtype.is_floating_point
        assert proxy_pos.dim() == 2 and proxy_pos.shape[0] == 3 and proxy_pos.is_contiguous()
        self._check(self._L.b200cv_read_data(self._h, C.c_void_p(proxy_pos.data_ptr()),
                                             proxy_pos.shape[1], _stream_ptr(stream)))

    def calc_value(self, step=0, gradients=True, stream=None):
        self._check(self._L.b200cv_calc_value(self._h, int(step), int(gradients),
                                              _stream_ptr(stream)))

    def step(self, proxy_pos, step=0, gradients=True, stream=None):
        """read_data + calc_value (+ fit gradients) replayed from a CUDA graph."""
        assert proxy_pos.is_cuda and proxy_pos.shape[0] == 3 and proxy_pos.is_contiguous()
        self._check(self._L.b200cv_step(self._h, C.c_void_p(proxy_pos.data_ptr()),
                                        proxy_pos.shape[1], int(step), int(gradients),
                                        _stream_ptr(stream)))

    def calc_gradients(self):
        """Gradients are produced by calc_value (src/colvarcomp_coordnums.cpp:316-319)."""

    def calc_fit_gradients(self, stream=None):
        self._check(self._L.b200cv_calc_fit_gradients(self._h, _stream_ptr(stream)))

    def value(self):
        v = C.c_double()
        self._check(self._L.b200cv_value(self._h, C.byref(v)))
        return v.value

    def apply_force(self, force, proxy_force, stream=None):
        assert proxy_force.is_cuda and proxy_force.shape[0] == 3 and proxy_force.is_contiguous()
        self._check(self._L.b200cv_apply_force(self._h, float(force),
                                               C.c_void_p(proxy_force.data_ptr()),
                                               proxy_force.shape[1], _stream_ptr(stream)))

    def eval(self, pos1, pos2=None, flags=EF_GRADIENTS, grad1=None, grad2=None, stream=None):
        """Bare-array entry: pos*/grad* are torch CUDA float64 tensors (3, n) (SoA)."""
        p2 = C.c_void_p(pos2.data_ptr()) if pos2 is not None else None
        g1 = C.c_void_p(grad1.data_ptr()) if grad1 is not None else None
        g2 = C.c_void_p(grad2.data_ptr()) if grad2 is not None else None
        self._check(self._L.b200cv_eval(self._h, C.c_void_p(pos1.data_ptr()), p2, int(flags), g1,
                                        g2, _stream_ptr(stream)))

    def eval_async(self, pos1, pos2=None, flags=EF_GRADIENTS, stream=None):
        """Capturable evaluation (no synchronisation, no allocation): the scalar is fetched
        with value() once the stream has drained, the gradients with accumulate_gradients()."""
        p2 = C.c_void_p(pos2.data_ptr()) if pos2 is not None else None
        self._check(self._L.b200cv_eval_async(self._h, C.c_void_p(pos1.data_ptr()), p2,
                                              int(flags), None, None, _stream_ptr(stream)))

    def accumulate_gradients(self, grad1, grad2=None, stream=None):
        g2 = C.c_void_p(grad2.data_ptr()) if grad2 is not None else None
        self._check(self._L.b200cv_accumulate_gradients(self._h, C.c_void_p(grad1.data_ptr()),
                                                        g2, _stream_ptr(stream)))

    def pairlist_reserve(self, pos1, pos2=None, headroom=2.0, stream=None):
        p2 = C.c_void_p(pos2.data_ptr()) if pos2 is not None else None
        self._check(self._L.b200cv_pairlist_reserve(self._h, C.c_void_p(pos1.data_ptr()), p2,
                                                    float(headroom), _stream_ptr(stream)))

    def pairlist_set_rebuild(self, rebuild):
        self._check(self._L.b200cv_pairlist_set_rebuild(self._h, int(bool(rebuild))))

    def async_set_skip(self, skip):
        self._check(self._L.b200cv_async_set_skip(self._h, int(bool(skip))))

    def eval_host(self, pos1, pos2=None, flags=EF_GRADIENTS, grad1=None, grad2=None):
        """pos*/grad*: numpy SoA arrays (3, n) float64 in host memory; returns the value."""
        v = C.c_double()
        self._check(self._L.b200cv_eval_host(self._h, _as_dp(pos1), _as_dp(pos2), int(flags),
                                             _as_dp(grad1), _as_dp(grad2), C.byref(v)))
        return v.value

    def eval_pairs_host(self, pos1, pos2, flags=EF_GRADIENTS, grad1=None, grad2=None):
        """Batch of single pairs (k, k); numpy SoA arrays (3, n); returns the sum."""
        v = C.c_double()
        self._check(self._L.b200cv_eval_pairs_host(self._h, _as_dp(pos1), _as_dp(pos2),
                                                   int(flags), _as_dp(grad1), _as_dp(grad2),
                                                   C.byref(v)))
        return v.value

    def distance_pairs_host(self, pos1, pos2):
        """distancePairs value: numpy SoA (3, n) positions -> (n1 * n2,) distances."""
        out = np.zeros(self.n1 * self.n2)
        self._check(self._L.b200cv_distance_pairs_host(self._h, _as_dp(pos1), _as_dp(pos2),
                                                       _as_dp(out)))
        return out

    def distance_pairs_force_host(self, force):
        """Forces on the atoms of both groups for a force vector on the distances."""
        f1, f2 = np.zeros(3 * self.n1), np.zeros(3 * self.n2)
        force = np.ascontiguousarray(force, dtype=np.float64)
        self._check(self._L.b200cv_distance_pairs_force_host(self._h, _as_dp(force), _as_dp(f1),
                                                             _as_dp(f2)))
        return f1.reshape(3, self.n1).T.copy(), f2.reshape(3, self.n2).T.copy()

    # -- host-buffer step (CPU proxy arrays, AoS) ---------------------------------------------
    def calc_host(self, proxy_pos_aos, step=0, gradients=True):
        """proxy_pos_aos: numpy (P, 3) float64, or a pinned torch CPU tensor of that shape."""
        ptr, n = _host_ptr(proxy_pos_aos)
        v = C.c_double()
        self._check(self._L.b200cv_calc_host(self._h, ptr, n, int(step), int(gradients),
                                             C.byref(v)))
        return v.value

    def apply_force_host(self, force, proxy_force_aos):
        ptr, n = _host_ptr(proxy_force_aos)
        self._check(self._L.b200cv_apply_force_host(self._h, float(force), ptr, n))

    # -- inspection ----------------------------------------------------------------------------
    def positions(self, which):
        n = self.n1 if which == 1 else self.n2
        out = np.zeros(3 * n)
        self._check(self._L.b200cv_get_positions(self._h, which, _as_dp(out)))
        return out.reshape(3, n).T.copy()

    def gradients(self, which):
        n = self.n1 if which == 1 else self.n2
        out = np.zeros(3 * n)
        self._check(self._L.b200cv_get_gradients(self._h, which, _as_dp(out)))
        return out.reshape(3, n).T.copy()

    def fit_gradients(self, which, nfit):
        out = np.zeros(3 * nfit)
        self._check(self._L.b200cv_get_fit_gradients(self._h, which, _as_dp(out), nfit))
        return out.reshape(3, nfit).T.copy()

    def centers(self, which):
        com, cog = np.zeros(3), np.zeros(3)
        self._check(self._L.b200cv_get_centers(self._h, which, _as_dp(com), _as_dp(cog)))
        return com, cog

    def rotation(self, which):
        q = np.zeros(4)
        self._check(self._L.b200cv_get_rotation(self._h, which, _as_dp(q)))
        return q

    def tolerance_l2_max(self):
        v = C.c_double()
        self._check(self._L.b200cv_get_tolerance_l2_max(self._h, C.byref(v)))
        return v.value

    def pairlist(self):
        """Sorted canonical indices of the listed pairs (uint64)."""
        n = C.c_size_t()
        self._check(self._L.b200cv_pairlist_count(self._h, C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=np.uint64)
        self._check(self._L.b200cv_pairlist_export(
            self._h, out.ctypes.data_as(C.POINTER(C.c_uint64)), out.size))
        return out[:n.value]

    def pairlist_stats(self):
        """dict(form, pairs, entries, row_entries, row_chunks); see b200cv_pairlist_stats."""
        form = C.c_int()
        v = [C.c_size_t() for _ in range(4)]
        self._check(self._L.b200cv_pairlist_stats(self._h, C.byref(form), *[C.byref(x) for x in v]))
        return dict(form=form.value, pairs=v[0].value, entries=v[1].value,
                    row_entries=v[2].value, row_chunks=v[3].value)

    def last_stats(self):
        k = C.c_int()
        e = C.c_size_t()
        self._check(self._L.b200cv_last_stats(self._h, C.byref(k), C.byref(e)))
        return k.value, e.value


def _host_ptr(a):
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float64 and a.flags.c_contiguous and a.ndim == 2 and a.shape[1] == 3
        return C.c_void_p(a.ctypes.data), a.shape[0]
    # torch CPU tensor (pinned or not)
    assert (not a.is_cuda) and a.is_contiguous() and a.dim() == 2 and a.shape[1] == 3
    return C.c_void_p(a.data_ptr()), a.shape[0]


def kernel_launches():
    """Kernels launched by libb200cv in this process so far."""
    return int(load().b200cv_kernel_launches())


def plan_items(function_type, n1, n2, rank=0, world=1):
    """Work items (i0, j0, jn, diag) this rank sweeps; pure host logic (no GPU needed)."""
    L = load()
    n = C.c_size_t()
    ch = C.c_int()
    kind = KIND_BY_NAME[function_type]
    rc = L.b200cv_plan_items(kind, n1, n2, rank, world, None, 0, C.byref(n), C.byref(ch))
    if rc != OK:
        raise B200cvError(rc, "plan_items")
    out = np.zeros((max(n.value, 1), 4), dtype=np.int32)
    rc = L.b200cv_plan_items(kind, n1, n2, rank, world, out.ctypes.data_as(_ip), n.value,
                             C.byref(n), C.byref(ch))
    if rc != OK:
        raise B200cvError(rc, "plan_items")
    return out[:n.value], ch.value


def comm_unique_id():
    buf = C.create_string_buffer(128)
    rc = load().b200cv_comm_unique_id(buf)
    if rc != OK:
        raise B200cvError(rc, load().b200cv_last_error(None).decode())
    return buf.raw


def fp64_peak(device=0, target_ms=200.0):
    t, ms = C.c_double(), C.c_double()
    rc = load().b200cv_fp64_peak(device, float(target_ms), C.byref(t), C.byref(ms))
    if rc != OK:
        raise B200cvError(rc, load().b200cv_last_error(None).decode())
    return t.value, ms.value


def selftest_rcp(device=0, n=1 << 20, lo=1.0, hi=1e30):
    u = C.c_double()
    rc = load().b200cv_selftest_rcp(device, n, lo, hi, C.byref(u))
    if rc != OK:
        raise B200cvError(rc, load().b200cv_last_error(None).decode())
    return u.value
