"""ctypes binding of the CPU oracle (oracle/nvsm_oracle.hpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, as the checker. Nothing under
cunvsm_b200/ imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

SGD, ADAGRAD, ADAM = 0, 1, 2
SPARSE, DENSE_UPDATE, DENSE_UPDATE_DENSE_VARIANCE = 1, 2, 3
TANH, HARD_TANH = 0, 1


def build(native=False):
    target = "native" if native else "all"
    subprocess.run(["make", "-C", _HERE, target], check=True, capture_output=True)
    return os.path.join(_HERE, "libnvsm_oracle_native.so" if native else "libnvsm_oracle.so")


class _Config(ctypes.Structure):
    _fields_ = [
        ("num_words", ctypes.c_long), ("num_entities", ctypes.c_long),
        ("word_repr_size", ctypes.c_int), ("entity_repr_size", ctypes.c_int),
        ("nonlinearity", ctypes.c_int), ("batch_normalization", ctypes.c_int),
        ("clip_sigmoid", ctypes.c_int), ("bias_negative_samples", ctypes.c_int),
        ("update_method", ctypes.c_int), ("adam_mode", ctypes.c_int),
        ("num_random_entities", ctypes.c_int),
        ("regularization_lambda", ctypes.c_double), ("bn_epsilon", ctypes.c_double),
    ]


def lib(native=False):
    key = bool(native)
    if key in _LIBS:
        return _LIBS[key]
    path = os.path.join(_HERE, "libnvsm_oracle_native.so" if native else "libnvsm_oracle.so")
    if not os.path.exists(path):
        path = build(native)
    L = ctypes.CDLL(path)
    vp, cl, ci, cd = ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_double
    pl = ctypes.POINTER(ctypes.c_long)
    pul = ctypes.POINTER(ctypes.c_ulong)
    L.oracle_generate_labels.argtypes = [pl, cl, cl, cl, pul, pl]
    L.oracle_generate_random_indexes.argtypes = [cl, cl, pul, pl]
    L.oracle_num_threads.restype = ci
    for suf, ct in (("f64", ctypes.c_double), ("f32", ctypes.c_float)):
        pf = ctypes.POINTER(ct)
        ppf = ctypes.POINTER(pf)
        ppl = ctypes.POINTER(pl)

        def f(name, argtypes, restype=None):
            fn = getattr(L, "oracle_%s_%s" % (name, suf))
            fn.argtypes = argtypes
            fn.restype = restype

        f("glorot", [pf, cl, cl, pul])
        f("truncated_sigmoid", [ct, ct], ct)
        f("sigmoid_deriv", [ct, ct], ct)
        f("clip", [ct], ct)
        f("clip_deriv", [ct], ct)
        f("gather_mean", [pf, cl, pl, pf, cl, cl, pf])
        f("update_dense", [pf, cl, pf, ct, ct, ci])
        f("bn_create", [cd], vp)
        f("bn_destroy", [vp])
        f("bn_forward", [vp, pf, pf, cl, cl, pf])
        f("bn_backward", [vp, pf, pf, cl, pf, pf])
        f("normalizer", [pf, pf, cl, cl, pf, pf])
        f("similarity_step", [pf, cl, pl, pf, cl, ci, pf, pf], ct)
        f("repr_updater_create", [ci, ci, cl, cl, ct, ct, ct], vp)
        f("repr_updater_destroy", [vp])
        f("repr_updater_update", [vp, pf, ci, ppf, ppl, pl, pl, ppf, ct, ct])
        f("repr_updater_state", [vp, ci, pf, cl], cl)
        f("transform_updater_create", [ci, cl, cl, ct, ct, ct], vp)
        f("transform_updater_destroy", [vp])
        f("transform_updater_update", [vp, pf, pf, pf, pf, ct, ct])
        f("transform_updater_state", [vp, ci, pf, cl], cl)
        f("model_create", [ctypes.POINTER(_Config)], vp)
        f("model_destroy", [vp])
        f("model_initialize", [vp, pul])
        f("model_array_size", [vp, ctypes.c_char_p], cl)
        f("model_get", [vp, ctypes.c_char_p, pf, cl], ci)
        f("model_set", [vp, ctypes.c_char_p, pf, cl], ci)
        f("model_compute_cost", [vp, pl, pf, pl, pf, cl, cl], ct)
        f("model_compute_gradients", [vp])
        f("model_update", [vp, ct, ct])
        f("model_scaled_lambda", [vp], ct)
        f("model_infer", [vp, pl, cl, cl, pf])
    _LIBS[key] = L
    return L


def _dt(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64", ctypes.c_double
    if dtype == np.float32:
        return "f32", ctypes.c_float
    raise TypeError(dtype)


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _pl(a):
    assert a.dtype == np.int64 and a.flags.c_contiguous
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_long))


def generate_labels(labels, z, num_objects, state):
    """cpp/labels.cu:3-22. Returns (ids[B*(z+1)], new_state)."""
    labels = np.ascontiguousarray(labels, dtype=np.int64)
    out = np.zeros(labels.size * (z + 1), dtype=np.int64)
    st = ctypes.c_ulong(state)
    lib().oracle_generate_labels(_pl(labels), labels.size, z, num_objects, ctypes.byref(st), _pl(out))
    return out, st.value


def generate_labels_cdf(labels, z, cdf, state):
    """Restatement of the inverse-CDF negative sampler (include/nvsm_b200.h: nvsm_generate_labels_cdf; the plug point is
    the reference's LabelGenerator, include/cuNVSM/labels.h:7-18). Per negative, in order: x <- 16807 x mod (2^31-1)
    (std::minstd_rand0), u = (x-1)/2147483646, id = #{k : cdf[k] <= u} clipped to D-1. Returns (ids[B*(z+1)], state)."""
    labels = np.asarray(labels, dtype=np.int64)
    cdf = np.asarray(cdf, dtype=np.float64)
    n = labels.size * z
    M = np.uint64(2147483647)
    # x_k = 16807^k x_0 mod M for k = 1..n; the powers by doubling (products stay below 2^62)
    pw = np.array([16807], dtype=np.uint64)
    while pw.size < n:
        pw = np.concatenate([pw, (pw * pw[-1]) % M])
    xs = ((pw[:n] * np.uint64(state)) % M).astype(np.int64)
    x = int(xs[-1]) if n else int(state)
    u = (xs - 1).astype(np.float64) / 2147483646.0
    neg = np.minimum(np.searchsorted(cdf, u, side="right"), cdf.size - 1).reshape(labels.size, z)
    out = np.concatenate([labels.reshape(-1, 1), neg], axis=1).astype(np.int64).ravel()
    return out, x


def generate_random_indexes(maxv, num, state):
    out = np.zeros(num, dtype=np.int64)
    st = ctypes.c_ulong(state)
    lib().oracle_generate_random_indexes(maxv, num, ctypes.byref(st), _pl(out))
    return out, st.value


def glorot(rows, cols, state, dtype=np.float64):
    suf, ct = _dt(dtype)
    out = np.zeros(rows * cols, dtype=dtype)
    st = ctypes.c_ulong(state)
    getattr(lib(), "oracle_glorot_" + suf)(_p(out, ct), rows, cols, ctypes.byref(st))
    return out, st.value


def scalar_fn(name, dtype, *args):
    suf, ct = _dt(dtype)
    return getattr(lib(), "oracle_%s_%s" % (name, suf))(*args)


def gather_mean(repr_, dim, idx, weights, window, dtype=np.float64):
    suf, ct = _dt(dtype)
    repr_ = np.ascontiguousarray(repr_, dtype=dtype)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    num_out = idx.size // window
    out = np.zeros((num_out, dim), dtype=dtype)
    w = None
    if weights is not None:
        weights = np.ascontiguousarray(weights, dtype=dtype)
        w = _p(weights, ct)
    getattr(lib(), "oracle_gather_mean_" + suf)(_p(repr_, ct), dim, _pl(idx), w, num_out, window, _p(out, ct))
    return out


def update_dense(param, grad, lr, lam, square=False):
    suf, ct = _dt(param.dtype)
    grad = np.ascontiguousarray(grad, dtype=param.dtype)
    getattr(lib(), "oracle_update_dense_" + suf)(_p(param, ct), param.size, _p(grad, ct), lr, lam, int(square))
    return param


def similarity_step(table, pair_ids, weights, clip_sigmoid=True, dtype=np.float64):
    """RepresentationSimilarity::Objective compute_cost + compute_gradients (cpp/objective.cu:485-672) over
    table[rows][dim]: returns (cost, probs[N], grad[2N][dim]) — grad rows belong to pair_ids.ravel(), window 1."""
    suf, ct = _dt(dtype)
    table = np.ascontiguousarray(table, dtype=dtype)
    ids = np.ascontiguousarray(pair_ids, dtype=np.int64).ravel()
    w = np.ascontiguousarray(weights, dtype=dtype)
    N = w.size
    assert ids.size == 2 * N
    probs = np.zeros(N, dtype=dtype)
    grad = np.zeros((2 * N, table.shape[1]), dtype=dtype)
    cost = getattr(lib(), "oracle_similarity_step_" + suf)(_p(table, ct), table.shape[1], _pl(ids), _p(w, ct), N,
                                                           int(clip_sigmoid), _p(probs, ct), _p(grad, ct))
    return float(cost), probs, grad


def normalizer(x, grad_output=None, dtype=np.float64):
    """Normalizer::forward (+ backward) of cpp/cuda_utils.cu:3-141 on instances x[N][dim]: (y, dx or None)."""
    suf, ct = _dt(dtype)
    x = np.ascontiguousarray(x, dtype=dtype)
    N, dim = x.shape
    y = np.zeros_like(x)
    dx = np.zeros_like(x) if grad_output is not None else None
    dy = np.ascontiguousarray(grad_output, dtype=dtype) if grad_output is not None else None
    getattr(lib(), "oracle_normalizer_" + suf)(_p(x, ct), _p(dy, ct) if dy is not None else None, N, dim, _p(y, ct),
                                               _p(dx, ct) if dx is not None else None)
    return y, dx


class BatchNorm:
    def __init__(self, eps, dtype=np.float64):
        self.suf, self.ct = _dt(dtype)
        self.dtype = dtype
        self.h = getattr(lib(), "oracle_bn_create_" + self.suf)(eps)

    def __del__(self):
        getattr(lib(), "oracle_bn_destroy_" + self.suf)(self.h)

    def forward(self, x, bias):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        bias = np.ascontiguousarray(bias, dtype=self.dtype)
        y = np.zeros_like(x)
        getattr(lib(), "oracle_bn_forward_" + self.suf)(self.h, _p(x, self.ct), _p(bias, self.ct), x.shape[0], x.shape[1], _p(y, self.ct))
        return y

    def backward(self, dy, x):
        dy = np.ascontiguousarray(dy, dtype=self.dtype)
        x = np.ascontiguousarray(x, dtype=self.dtype)
        dx = np.zeros_like(dy)
        db = np.zeros(dy.shape[1], dtype=self.dtype)
        getattr(lib(), "oracle_bn_backward_" + self.suf)(self.h, _p(dy, self.ct), _p(x, self.ct), dy.shape[0], _p(dx, self.ct), _p(db, self.ct))
        return dx, db


class ReprUpdater:
    """Free-standing table optimiser (cpp/updates*.cu RepresentationsGradientUpdater)."""

    def __init__(self, method, adam_mode, num_objects, dim, beta1=0.9, beta2=0.999, eps=1e-6, dtype=np.float64):
        self.suf, self.ct = _dt(dtype)
        self.dtype = dtype
        self.num_objects, self.dim = num_objects, dim
        self.h = getattr(lib(), "oracle_repr_updater_create_" + self.suf)(method, adam_mode, num_objects, dim, beta1, beta2, eps)

    def __del__(self):
        getattr(lib(), "oracle_repr_updater_destroy_" + self.suf)(self.h)

    def update(self, table, descs, lr, lam):
        """descs: list of (grad[num_grads, dim] (modified in place), idx, window, weights|None)."""
        n = len(descs)
        pf = ctypes.POINTER(self.ct)
        pl = ctypes.POINTER(ctypes.c_long)
        grads = (pf * n)(); idxs = (pl * n)(); wts = (pf * n)()
        ng = (ctypes.c_long * n)(); win = (ctypes.c_long * n)()
        keep = []
        for i, (g, idx, window, w) in enumerate(descs):
            assert g.dtype == self.dtype and g.flags.c_contiguous
            idx = np.ascontiguousarray(idx, dtype=np.int64)
            keep.append(idx)
            grads[i] = _p(g, self.ct); idxs[i] = _pl(idx)
            ng[i] = g.shape[0]; win[i] = window
            if w is not None:
                w = np.ascontiguousarray(w, dtype=self.dtype)
                keep.append(w)
                wts[i] = _p(w, self.ct)
            else:
                wts[i] = None
        getattr(lib(), "oracle_repr_updater_update_" + self.suf)(self.h, _p(table, self.ct), n, grads, idxs, ng, win, wts, lr, lam)

    def state(self, which):
        fn = getattr(lib(), "oracle_repr_updater_state_" + self.suf)
        n = fn(self.h, which, None, 0)
        out = np.zeros(n, dtype=self.dtype)
        fn(self.h, which, _p(out, self.ct), n)
        return out


class TransformUpdater:
    def __init__(self, method, nT, nb, beta1=0.9, beta2=0.999, eps=1e-6, dtype=np.float64):
        self.suf, self.ct = _dt(dtype)
        self.dtype = dtype
        self.h = getattr(lib(), "oracle_transform_updater_create_" + self.suf)(method, nT, nb, beta1, beta2, eps)

    def __del__(self):
        getattr(lib(), "oracle_transform_updater_destroy_" + self.suf)(self.h)

    def update(self, T, b, gT, gb, lr, lam):
        getattr(lib(), "oracle_transform_updater_update_" + self.suf)(self.h, _p(T, self.ct), _p(b, self.ct), _p(gT, self.ct), _p(gb, self.ct), lr, lam)

    def state(self, which):
        fn = getattr(lib(), "oracle_transform_updater_state_" + self.suf)
        n = fn(self.h, which, None, 0)
        out = np.zeros(n, dtype=self.dtype)
        fn(self.h, which, _p(out, self.ct), n)
        return out


class Model:
    """The oracle TextEntity model: compute_cost / compute_gradients / update."""

    def __init__(self, num_words, num_entities, word_repr_size, entity_repr_size, *, nonlinearity=TANH,
                 batch_normalization=False, clip_sigmoid=False, bias_negative_samples=False,
                 update_method=SGD, adam_mode=SPARSE, num_random_entities=1, regularization_lambda=0.0,
                 bn_epsilon=1e-4, dtype=np.float64, native=False):
        self.suf, self.ct = _dt(dtype)
        self.dtype = np.dtype(dtype)
        self.L = lib(native)
        self.cfg = _Config(num_words, num_entities, word_repr_size, entity_repr_size, nonlinearity,
                           int(batch_normalization), int(clip_sigmoid), int(bias_negative_samples),
                           update_method, adam_mode, num_random_entities, regularization_lambda, bn_epsilon)
        self.h = self._fn("model_create")(ctypes.byref(self.cfg))

    def _fn(self, name):
        return getattr(self.L, "oracle_%s_%s" % (name, self.suf))

    def __del__(self):
        try:
            self._fn("model_destroy")(self.h)
        except Exception:
            pass

    def initialize(self, state):
        st = ctypes.c_ulong(state)
        self._fn("model_initialize")(self.h, ctypes.byref(st))
        return st.value

    def get(self, name):
        n = self._fn("model_array_size")(self.h, name.encode())
        if n < 0:
            raise KeyError(name)
        out = np.zeros(n, dtype=self.dtype)
        assert self._fn("model_get")(self.h, name.encode(), _p(out, self.ct), n) == 0
        return out

    def set(self, name, value):
        value = np.ascontiguousarray(value, dtype=self.dtype).ravel()
        rc = self._fn("model_set")(self.h, name.encode(), _p(value, self.ct), value.size)
        if rc != 0:
            raise ValueError("bad size for %s" % name)

    def compute_cost(self, features, feature_weights, entity_ids, weights, window):
        features = np.ascontiguousarray(features, dtype=np.int64).ravel()
        fw = np.ascontiguousarray(feature_weights, dtype=self.dtype).ravel()
        ids = np.ascontiguousarray(entity_ids, dtype=np.int64).ravel()
        w = np.ascontiguousarray(weights, dtype=self.dtype).ravel()
        B = w.size
        assert features.size == B * window and ids.size == B * (self.cfg.num_random_entities + 1)
        return self._fn("model_compute_cost")(self.h, _pl(features), _p(fw, self.ct), _pl(ids), _p(w, self.ct), B, window)

    def compute_gradients(self):
        self._fn("model_compute_gradients")(self.h)

    def update(self, lr, scaled_lambda):
        self._fn("model_update")(self.h, lr, scaled_lambda)

    def scaled_lambda(self):
        return self._fn("model_scaled_lambda")(self.h)

    def infer(self, words, window):
        words = np.ascontiguousarray(words, dtype=np.int64).ravel()
        N = words.size // window
        out = np.zeros((N, self.cfg.entity_repr_size), dtype=self.dtype)
        self._fn("model_infer")(self.h, _pl(words), N, window, _p(out, self.ct))
        return out
