"""ctypes binding of libcpd_b200.so (include/cpd_b200.h).  No torch, no cupy, no CPU fallback:
if the shared library is missing or no CUDA device is visible, using the package raises."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CPD_B200_LIB: load another build of the same library (tools/tune.sh variants); never a different implementation -- lib() refuses
# anything that exports cpd_is_emulation (the CPU build the test-suite keeps for itself), so the package cannot be given a CPU path this way
LIB_PATH = os.environ.get("CPD_B200_LIB") or os.path.join(_HERE, "libcpd_b200.so")

TF_RIGID, TF_AFFINE, TF_NONRIGID = 0, 1, 2


class CpdParams(ctypes.Structure):
    _fields_ = [("lin", ctypes.c_double * 9), ("t", ctypes.c_double * 3), ("scale", ctypes.c_double),
                ("sigma2", ctypes.c_double), ("q", ctypes.c_double), ("n_p", ctypes.c_double)]


class CpdError(RuntimeError):
    pass


_c_dp = ctypes.POINTER(ctypes.c_double)
_c_fp = ctypes.POINTER(ctypes.c_float)
_PROTOS = {
    "cpd_last_error": (ctypes.c_char_p, []),
    "cpd_version": (ctypes.c_int, []),
    "cpd_device_count": (ctypes.c_int, []),
    "cpd_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "cpd_destroy": (None, [ctypes.c_void_p]),
    "cpd_set_source": (ctypes.c_int, [ctypes.c_void_p, _c_dp, ctypes.c_int64]),
    "cpd_set_target": (ctypes.c_int, [ctypes.c_void_p, _c_dp, ctypes.c_int64, ctypes.c_int64, _c_dp]),
    "cpd_sigma2_init": (ctypes.c_int, [ctypes.c_void_p, _c_dp]),
    "cpd_set_state": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.POINTER(CpdParams)]),
    "cpd_em_step": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(CpdParams)]),
    "cpd_em_run": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.POINTER(CpdParams),
                                  ctypes.POINTER(ctypes.c_int), _c_dp]),
    "cpd_estep": (ctypes.c_int, [ctypes.c_void_p, _c_dp, ctypes.c_double, ctypes.c_double, _c_dp, _c_dp, _c_dp, _c_dp]),
    "cpd_mstep": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _c_dp, _c_dp, _c_dp, ctypes.c_double,
                                 ctypes.POINTER(CpdParams)]),
    "cpd_bcpd_estep": (ctypes.c_int, [ctypes.c_void_p, _c_dp, ctypes.c_double, _c_dp, _c_dp, ctypes.c_double, ctypes.c_double,
                                      _c_dp, _c_dp, _c_dp, _c_dp]),
    "cpd_last_estep": (ctypes.c_int, [ctypes.c_void_p, _c_dp, _c_dp, _c_dp, _c_dp]),
    "cpd_nonrigid_begin": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double]),
    "cpd_nonrigid_step": (ctypes.c_int, [ctypes.c_void_p, _c_dp]),
    "cpd_nonrigid_get": (ctypes.c_int, [ctypes.c_void_p, _c_dp, _c_dp]),
    "cpd_nonrigid_restart": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_double]),
    "cpd_nonrigid_mstep": (ctypes.c_int, [ctypes.c_void_p, _c_dp, _c_dp, _c_dp, ctypes.c_double, _c_dp]),
    "cpd_nonrigid_lowrank_begin": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                                  ctypes.c_int, ctypes.c_int, ctypes.c_uint64]),
    "cpd_nonrigid_lowrank_get": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), _c_dp, _c_dp]),
    "cpd_nonrigid_set_prior": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, _c_dp, _c_dp]),
    "cpd_rbf_kernel": (ctypes.c_int, [ctypes.c_int, _c_dp, ctypes.c_int64, _c_dp, ctypes.c_int64, ctypes.c_int,
                                      ctypes.c_double, _c_fp]),
    "cpd_imq_kernel": (ctypes.c_int, [ctypes.c_int, _c_dp, ctypes.c_int64, _c_dp, ctypes.c_int64, ctypes.c_int,
                                      ctypes.c_double, _c_fp]),
    "cpd_gauss_transform": (ctypes.c_int, [ctypes.c_int, _c_dp, ctypes.c_int64, _c_dp, ctypes.c_int64, ctypes.c_int, ctypes.c_double,
                                           _c_dp, ctypes.c_int, _c_dp]),
    "cpd_squared_kernel_sum": (ctypes.c_int, [ctypes.c_int, _c_dp, ctypes.c_int64, _c_dp, ctypes.c_int64, ctypes.c_int, _c_dp]),
    "cpd_comm_unique_id": (ctypes.c_int, [ctypes.c_char_p]),
    "cpd_comm_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p]),
    "cpd_comm_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "cpd_comm_attach": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "cpd_p2p_local_handle": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p]),
    "cpd_p2p_attach": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]),
    "cpd_plan_work": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "cpd_p2p_detach": (ctypes.c_int, [ctypes.c_void_p]),
    "cpd_timer_start": (ctypes.c_int, [ctypes.c_void_p]),
    "cpd_timer_stop": (ctypes.c_int, [ctypes.c_void_p, _c_fp]),
    "cpd_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "cpd_event_record": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "cpd_event_elapsed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _c_fp]),
    "cpd_set_profiling": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "cpd_stage_times": (ctypes.c_int, [ctypes.c_void_p, _c_fp]),
    "cpd_lowrank_setup_times": (ctypes.c_int, [ctypes.c_void_p, _c_fp]),
    "cpd_launch_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "cpd_flush_l2": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64]),
    "cpd_microbench": (ctypes.c_int, [ctypes.c_int, _c_dp]),
}
EXPORTED = tuple(_PROTOS)
_lib = None


def _load(path):
    """dlopen `path` and attach the prototypes of include/cpd_b200.h."""
    handle = ctypes.CDLL(path)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    return handle


def lib():
    """The loaded shared library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CpdError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(or `make -C probreg_b200/csrc`); probreg_b200 has no CPU path" % LIB_PATH)
        loaded = _load(LIB_PATH)
        if hasattr(loaded, "cpd_is_emulation"):
            raise CpdError("%s is the CPU emulation used by the tests (it exports cpd_is_emulation): probreg_b200 has no CPU path "
                           "and will not load it as the product library" % LIB_PATH)
        _lib = loaded
    return _lib


def check(code):
    if code != 0:
        raise CpdError("libcpd_b200: %s (code %d)" % (lib().cpd_last_error().decode(), code))


def as_cloud(a, dim=None):
    """C-order float64 (count x D) view/copy of `a` (what cv() at probreg/cpd.py:444 hands on)."""
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    assert arr.ndim == 2, "source and target must have 2 dimensions."
    if dim is not None and arr.shape[1] != dim:
        raise ValueError("expected %d-D points, got %d-D" % (dim, arr.shape[1]))
    return arr


def dptr(a):
    return a.ctypes.data_as(_c_dp) if a is not None else None


class Handle(object):
    """RAII wrapper of cpd_ctx*: one GPU, one stream."""

    def __init__(self, dim, device=0, stream=None):
        if dim not in (2, 3):
            raise ValueError("probreg_b200 supports 2-D and 3-D points, got %d-D" % dim)
        self._h = ctypes.c_void_p()
        self.dim = dim
        self.device = device
        self._lib = lib()          # a handle is destroyed by the library that created it
        check(self._lib.cpd_create(ctypes.byref(self._h), device, dim, ctypes.c_void_p(stream) if stream else None))
        self.m = 0
        self.n = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.cpd_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- data
    def set_source(self, source):
        src = as_cloud(source, self.dim)
        check(self._lib.cpd_set_source(self._h, dptr(src), src.shape[0]))
        self.m = src.shape[0]

    def set_target(self, target, n_global=None, frame_origin=None):
        tgt = as_cloud(target, self.dim)
        n_global = tgt.shape[0] if n_global is None else int(n_global)
        org = None if frame_origin is None else np.ascontiguousarray(frame_origin, dtype=np.float64)
        check(self._lib.cpd_set_target(self._h, dptr(tgt), tgt.shape[0], n_global, dptr(org)))
        self.n = tgt.shape[0]

    def sigma2_init(self):
        out = ctypes.c_double()
        check(self._lib.cpd_sigma2_init(self._h, ctypes.byref(out)))
        return out.value

    # -- EM
    def set_state(self, tf_kind, update_scale, w, lin, t, scale, sigma2, q):
        p = CpdParams()
        d = self.dim
        lin = np.asarray(lin, dtype=np.float64).reshape(d, d)
        for i in range(d):
            for j in range(d):
                p.lin[i * d + j] = lin[i, j]
        tt = np.asarray(t, dtype=np.float64).reshape(d)
        for i in range(d):
            p.t[i] = tt[i]
        p.scale, p.sigma2, p.q = float(scale), float(sigma2), float(q)
        check(self._lib.cpd_set_state(self._h, tf_kind, int(bool(update_scale)), float(w), ctypes.byref(p)))

    def _unpack(self, p):
        d = self.dim
        lin = np.array(p.lin[: d * d], dtype=np.float64).reshape(d, d)
        t = np.array(p.t[:d], dtype=np.float64)
        return lin, t, p.scale, p.sigma2, p.q, p.n_p

    def em_step(self, read=True):
        if not read:
            check(self._lib.cpd_em_step(self._h, None))
            return None
        p = CpdParams()
        check(self._lib.cpd_em_step(self._h, ctypes.byref(p)))
        return self._unpack(p)

    def em_run(self, maxiter, tol, trace=False):
        p = CpdParams()
        it = ctypes.c_int()
        tr = np.zeros((max(maxiter, 1), 2)) if trace else None
        check(self._lib.cpd_em_run(self._h, int(maxiter), float(tol), ctypes.byref(p), ctypes.byref(it), dptr(tr)))
        out = self._unpack(p) + (it.value,)
        return out + (tr[: it.value],) if trace else out

    def estep(self, t_source, sigma2, w, want_pt1=True, want_p1=True, want_px=True):
        ts = as_cloud(t_source, self.dim)
        if ts.shape[0] != self.m:
            raise ValueError("t_source has %d rows, the handle's source has %d" % (ts.shape[0], self.m))
        pt1 = np.empty(self.n) if want_pt1 else None
        p1 = np.empty(self.m) if want_p1 else None
        px = np.empty((self.m, self.dim)) if want_px else None
        n_p = ctypes.c_double()
        check(self._lib.cpd_estep(self._h, dptr(ts), float(sigma2), float(w), dptr(pt1), dptr(p1), dptr(px), ctypes.byref(n_p)))
        return pt1, p1, px, n_p.value

    def bcpd_estep(self, t_source, scale, alpha, sigma_diag, sigma2, w):
        """(nu_d, nu, px, n_p) of probreg/bcpd.py:53-72 for the handle's target."""
        ts = as_cloud(t_source, self.dim)
        al = np.ascontiguousarray(alpha, dtype=np.float64)
        sd = np.ascontiguousarray(sigma_diag, dtype=np.float64)
        if ts.shape[0] != self.m or al.shape != (self.m,) or sd.shape != (self.m,):
            raise ValueError("t_source / alpha / sigma_diag do not match the handle's source count %d" % self.m)
        nu_d, nu, px = np.empty(self.n), np.empty(self.m), np.empty((self.m, self.dim))
        n_p = ctypes.c_double()
        check(self._lib.cpd_bcpd_estep(self._h, dptr(ts), float(scale), dptr(al), dptr(sd), float(sigma2), float(w), dptr(nu_d), dptr(nu),
                                   dptr(px), ctypes.byref(n_p)))
        return nu_d, nu, px, n_p.value

    def last_estep(self):
        pt1, p1, px = np.empty(self.n), np.empty(self.m), np.empty((self.m, self.dim))
        n_p = ctypes.c_double()
        check(self._lib.cpd_last_estep(self._h, dptr(pt1), dptr(p1), dptr(px), ctypes.byref(n_p)))
        return pt1, p1, px, n_p.value

    def mstep(self, tf_kind, update_scale, pt1, p1, px, n_p):
        pt1 = np.ascontiguousarray(pt1, dtype=np.float64)
        p1 = np.ascontiguousarray(p1, dtype=np.float64)
        px = as_cloud(px, self.dim)
        if pt1.shape[0] != self.n or p1.shape[0] != self.m or px.shape[0] != self.m:
            raise ValueError("EstepResult shapes do not match the handle's source/target")
        p = CpdParams()
        check(self._lib.cpd_mstep(self._h, tf_kind, int(bool(update_scale)), dptr(pt1), dptr(p1), dptr(px), float(n_p),
                              ctypes.byref(p)))
        return self._unpack(p)

    # -- non-rigid (dense G on the device)
    def nonrigid_begin(self, beta, lmd, sigma2, w):
        check(self._lib.cpd_nonrigid_begin(self._h, float(beta), float(lmd), float(sigma2), float(w)))

    def nonrigid_lowrank_begin(self, beta, lmd, sigma2, w, rank, power_iters=2, seed=0):
        check(self._lib.cpd_nonrigid_lowrank_begin(self._h, float(beta), float(lmd), float(sigma2), float(w), int(rank), int(power_iters),
                                               int(seed)))

    def nonrigid_lowrank_factors(self):
        """(Q (m x rank), Bc (rank x rank)) with G ~= Q Bc Q^T, Q in the caller's point order."""
        k = ctypes.c_int()
        check(self._lib.cpd_nonrigid_lowrank_get(self._h, ctypes.byref(k), None, None))
        q, b = np.empty((self.m, k.value)), np.empty((k.value, k.value))
        check(self._lib.cpd_nonrigid_lowrank_get(self._h, None, dptr(q), dptr(b)))
        return q, b

    def nonrigid_set_prior(self, alpha, p1_tilde, px_tilde):
        if p1_tilde is None:
            check(self._lib.cpd_nonrigid_set_prior(self._h, 1.0, None, None))
            return
        p1t = np.ascontiguousarray(p1_tilde, dtype=np.float64)
        pxt = as_cloud(px_tilde, self.dim)
        if p1t.shape != (self.m,) or pxt.shape[0] != self.m:
            raise ValueError("prior shapes do not match the handle's source")
        check(self._lib.cpd_nonrigid_set_prior(self._h, float(alpha), dptr(p1t), dptr(pxt)))

    def nonrigid_moved(self):
        t = np.empty((self.m, self.dim))
        check(self._lib.cpd_nonrigid_get(self._h, None, dptr(t)))
        return t

    def nonrigid_restart(self, lmd, sigma2, w):
        check(self._lib.cpd_nonrigid_restart(self._h, float(lmd), float(sigma2), float(w)))

    def nonrigid_mstep(self, pt1, p1, px, sigma2_p):
        pt1 = np.ascontiguousarray(pt1, dtype=np.float64)
        p1 = np.ascontiguousarray(p1, dtype=np.float64)
        px = as_cloud(px, self.dim)
        if pt1.shape[0] != self.n or p1.shape[0] != self.m or px.shape[0] != self.m:
            raise ValueError("EstepResult shapes do not match the handle's source/target")
        out = ctypes.c_double()
        check(self._lib.cpd_nonrigid_mstep(self._h, dptr(pt1), dptr(p1), dptr(px), float(sigma2_p), ctypes.byref(out)))
        return out.value

    def nonrigid_step(self):
        out = ctypes.c_double()
        check(self._lib.cpd_nonrigid_step(self._h, ctypes.byref(out)))
        return out.value

    def nonrigid_w(self):
        w = np.empty((self.m, self.dim))
        check(self._lib.cpd_nonrigid_get(self._h, dptr(w), None))
        return w

    # -- multi-GPU
    def attach_comm(self, nccl_comm, world_size, rank):
        """nccl_comm: the value returned by comm_create (borrowed; it must outlive the handle)."""
        check(self._lib.cpd_comm_attach(self._h, nccl_comm, world_size, rank))

    def p2p_local_handle(self):
        buf = ctypes.create_string_buffer(64)
        check(self._lib.cpd_p2p_local_handle(self._h, buf))
        return buf.raw

    def p2p_attach(self, handles, world_size, rank):
        blob = b"".join(handles)
        assert len(blob) == 64 * world_size
        check(self._lib.cpd_p2p_attach(self._h, blob, world_size, rank))

    def p2p_detach(self):
        check(self._lib.cpd_p2p_detach(self._h))

    # -- measurement
    def timer_start(self):
        check(self._lib.cpd_timer_start(self._h))

    def timer_stop(self):
        ms = ctypes.c_float()
        check(self._lib.cpd_timer_stop(self._h, ctypes.byref(ms)))
        return ms.value

    def sync(self):
        check(self._lib.cpd_sync(self._h))

    def event_record(self, idx):
        check(self._lib.cpd_event_record(self._h, idx))

    def event_elapsed(self, a, b):
        ms = ctypes.c_float()
        check(self._lib.cpd_event_elapsed(self._h, a, b, ctypes.byref(ms)))
        return ms.value

    def set_profiling(self, on):
        check(self._lib.cpd_set_profiling(self._h, int(on)))

    def stage_times(self):
        ms = (ctypes.c_float * 6)()
        check(self._lib.cpd_stage_times(self._h, ms))
        return list(ms)

    def lowrank_setup_times(self):
        ms = (ctypes.c_float * 3)()
        check(self._lib.cpd_lowrank_setup_times(self._h, ms))
        return {"gram_products_ms": ms[0], "orthonormalisation_ms": ms[1], "core_ms": ms[2]}

    def launch_count(self):
        return int(lib().cpd_launch_count(self._h))

    def flush_l2(self, nbytes=0):
        check(self._lib.cpd_flush_l2(self._h, nbytes))


def unique_id():
    buf = ctypes.create_string_buffer(128)
    check(lib().cpd_comm_unique_id(buf))
    return buf.raw


def plan_work(ntiles, nunits, slots, last_tile_cost=1.0):
    """(items (k x 4 int array: tile, first unit, end unit, slot), max partial slots per tile) -- host only."""
    n, mx = ctypes.c_int(), ctypes.c_int()
    check(lib().cpd_plan_work(ntiles, nunits, slots, float(last_tile_cost), None, 0, ctypes.byref(n), ctypes.byref(mx)))
    buf = np.zeros((n.value, 4), dtype=np.int32)
    check(lib().cpd_plan_work(ntiles, nunits, slots, float(last_tile_cost), buf.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), n.value,
                              ctypes.byref(n), ctypes.byref(mx)))
    return buf, mx.value


def comm_create(device, world_size, rank, uid):
    """Collective: every rank calls it once with the same unique id.  Returns an opaque pointer."""
    c = ctypes.c_void_p()
    check(lib().cpd_comm_create(ctypes.byref(c), device, world_size, rank, uid))
    return c


def comm_destroy(c):
    check(lib().cpd_comm_destroy(c))


def microbench(device=0):
    out = (ctypes.c_double * 9)()
    check(lib().cpd_microbench(device, out))
    return {"ffma_tflops": out[0], "mufu_ex2_gops": out[1], "sm_mhz": out[2], "sm_count": int(out[3]),
            "ffma2_tflops": out[4], "mix_11p1_gpairs": out[5], "mix_packed_gpairs": out[6], "mix_7p1_gpairs": out[7], "ffma2_plus_ffma_tflops": out[8]}
