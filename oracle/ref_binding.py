"""ctypes binding of oracle/_ref/libcunvsm_ref_{f32,f64}.so: the UNMODIFIED reference training step
(`Model<TextEntity::Objective>` of /root/reference) compiled by oracle/ref_shim/Makefile against a
reconstruction of its un-vendored device_matrix dependency.

TEST INFRASTRUCTURE ONLY: imported by tests/ and by bench.py's `--impl reference` arm. Needs a GPU: the
library creates a cuDNN handle when it is loaded (cpp/cudnn_utils.cu:186-187 of the reference).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

SGD, ADAGRAD, ADAM = 0, 1, 2
SPARSE, DENSE_UPDATE, DENSE_UPDATE_DENSE_VARIANCE = 1, 2, 3
TANH, HARD_TANH = 0, 1


def lib_path(dtype):
    return os.path.join(_HERE, "_ref", "libcunvsm_ref_%s.so" % ("f32" if np.dtype(dtype) == np.float32 else "f64"))


def available(dtype=np.float32):
    return os.path.exists(lib_path(dtype))


def lib(dtype):
    dtype = np.dtype(dtype)
    if dtype in _LIBS:
        return _LIBS[dtype]
    L = ctypes.CDLL(lib_path(dtype))
    ct = ctypes.c_float if dtype == np.float32 else ctypes.c_double
    vp, cl, ci, cd, cul = ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_double, ctypes.c_ulong
    pf, pl = ctypes.POINTER(ct), ctypes.POINTER(ctypes.c_long)
    sig = {
        "ref_float_bytes": ([], ci),
        "ref_create": ([cl, cl, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, cd, cul], vp),
        "ref_destroy": ([vp], None),
        "ref_get_rng_state": ([vp], cul),
        "ref_set_rng_state": ([vp, cul], None),
        "ref_tensor_size": ([vp, ctypes.c_char_p], cl),
        "ref_get_tensor": ([vp, ctypes.c_char_p, pf], ci),
        "ref_set_tensor": ([vp, ctypes.c_char_p, pf], ci),
        "ref_get_entity_ids": ([vp, pl, cl], cl),
        "ref_batch_create": ([cl, cl], vp),
        "ref_batch_destroy": ([vp], None),
        "ref_batch_fill": ([vp, pl, pf, pl, pf, cl], None),
        "ref_forward": ([vp, vp], None),
        "ref_get_cost": ([vp], cd),
        "ref_scaled_regularization_lambda": ([vp], cd),
        "ref_compute_gradients": ([vp], None),
        "ref_update": ([vp, cd, cd], None),
        "ref_step": ([vp, vp, cd], cd),
        "ref_synchronize": ([], None),
        "ref2_create": ([ci, cl, cl, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, cd, cd, cd, cul], vp),
        "ref2_destroy": ([vp], None),
        "ref2_get_rng_state": ([vp], cul),
        "ref2_tensor_size": ([vp, ctypes.c_char_p], cl),
        "ref2_get_tensor": ([vp, ctypes.c_char_p, pf], ci),
        "ref2_fill_text": ([vp, pl, pf, pl, pf, cl], None),
        "ref2_fill_pairs": ([vp, pl, pf, cl], None),
        "ref2_forward": ([vp], None),
        "ref2_get_cost": ([vp], cd),
        "ref2_scaled_regularization_lambda": ([vp], cd),
        "ref2_compute_gradients": ([vp], None),
        "ref2_update": ([vp, cd, cd], None),
    }
    for name, (argtypes, restype) in sig.items():
        fn = getattr(L, name)
        fn.argtypes, fn.restype = argtypes, restype
    assert L.ref_float_bytes() == dtype.itemsize
    L._ct = ct
    _LIBS[dtype] = L
    return L


class Batch:
    """TextEntity::Batch of the reference (pinned host arrays)."""

    def __init__(self, L, batch_size, window_size, dtype):
        self.L, self.dtype, self.window = L, np.dtype(dtype), window_size
        self.h = L.ref_batch_create(batch_size, window_size)

    def fill(self, features, labels, feature_weights, weights):
        f = np.ascontiguousarray(features, dtype=np.int64).ravel()
        fw = np.ascontiguousarray(feature_weights, dtype=self.dtype).ravel()
        lab = np.ascontiguousarray(labels, dtype=np.int64).ravel()
        w = np.ascontiguousarray(weights, dtype=self.dtype).ravel()
        assert f.size == lab.size * self.window == fw.size and w.size == lab.size
        pl, pf = ctypes.POINTER(ctypes.c_long), ctypes.POINTER(self.L._ct)
        self.L.ref_batch_fill(self.h, f.ctypes.data_as(pl), fw.ctypes.data_as(pf), lab.ctypes.data_as(pl),
                              w.ctypes.data_as(pf), lab.size)
        return self

    def __del__(self):
        try:
            self.L.ref_batch_destroy(self.h)
        except Exception:
            pass


class Model:
    """The reference's Model<TextEntity::Objective>; construction also runs Model::initialize(rng(seed))."""

    def __init__(self, num_words, num_entities, word_repr_size, entity_repr_size, *, batch_size, window_size,
                 num_random_entities, nonlinearity=TANH, batch_normalization=False, clip_sigmoid=False,
                 bias_negative_samples=False, l2_normalize_phrase_reprs=False, l2_normalize_entity_reprs=False,
                 update_method=SGD, adam_mode=SPARSE, regularization_lambda=0.0, seed=1, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.L = lib(self.dtype)
        self.shape = dict(V=num_words, D=num_entities, dw=word_repr_size, dd=entity_repr_size,
                          B=batch_size, n=window_size, R=num_random_entities + 1)
        self.h = self.L.ref_create(num_words, num_entities, word_repr_size, entity_repr_size, nonlinearity,
                                   int(batch_normalization), int(clip_sigmoid), int(bias_negative_samples),
                                   int(l2_normalize_phrase_reprs), int(l2_normalize_entity_reprs),
                                   update_method, adam_mode, batch_size, window_size, num_random_entities,
                                   float(regularization_lambda), seed)

    def __del__(self):
        try:
            self.L.ref_destroy(self.h)
        except Exception:
            pass

    def new_batch(self):
        return Batch(self.L, self.shape["B"], self.shape["n"], self.dtype)

    @property
    def rng_state(self):
        return self.L.ref_get_rng_state(self.h)

    @rng_state.setter
    def rng_state(self, state):
        self.L.ref_set_rng_state(self.h, state)

    def get(self, name):
        n = self.L.ref_tensor_size(self.h, name.encode())
        if n < 0:
            raise KeyError(name)
        out = np.zeros(n, dtype=self.dtype)
        assert self.L.ref_get_tensor(self.h, name.encode(), out.ctypes.data_as(ctypes.POINTER(self.L._ct))) == 0
        return out

    def set(self, name, value):
        value = np.ascontiguousarray(value, dtype=self.dtype).ravel()
        assert value.size == self.L.ref_tensor_size(self.h, name.encode()), name
        assert self.L.ref_set_tensor(self.h, name.encode(), value.ctypes.data_as(ctypes.POINTER(self.L._ct))) == 0

    def entity_ids(self):
        out = np.zeros(self.shape["B"] * self.shape["R"], dtype=np.int64)
        n = self.L.ref_get_entity_ids(self.h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_long)), out.size)
        assert n >= 0
        return out[:n]

    def forward(self, batch):
        self.L.ref_forward(self.h, batch.h)

    def get_cost(self):
        return self.L.ref_get_cost(self.h)

    def scaled_lambda(self):
        return self.L.ref_scaled_regularization_lambda(self.h)

    def compute_gradients(self):
        self.L.ref_compute_gradients(self.h)

    def update(self, lr, scaled_lambda):
        self.L.ref_update(self.h, lr, scaled_lambda)

    def step(self, batch, lr):
        return self.L.ref_step(self.h, batch.h, lr)

    def synchronize(self):
        self.L.ref_synchronize()


ENTITY_ENTITY, TERM_TERM, TEXT_ENTITY_ENTITY_ENTITY, TEXT_ENTITY_TERM_TERM = 1, 2, 3, 4


class ObjectiveModel:
    """Model<EntityEntity / TermTerm / TextEntityEntityEntity / TextEntityTermTerm ::Objective> of the reference
    (cpp/model.cu:222-228); owns its batch(es)."""

    def __init__(self, objective, num_words, num_entities, word_repr_size, entity_repr_size, *, batch_size, window_size,
                 num_random_entities, similarity_batch_size, nonlinearity=TANH, batch_normalization=False,
                 clip_sigmoid=False, bias_negative_samples=False, update_method=SGD, adam_mode=SPARSE,
                 regularization_lambda=0.0, text_entity_weight=1.0, similarity_weight=0.0, seed=1, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.L = lib(self.dtype)
        self.window = window_size
        self.h = self.L.ref2_create(objective, num_words, num_entities, word_repr_size, entity_repr_size, nonlinearity,
                                    int(batch_normalization), int(clip_sigmoid), int(bias_negative_samples),
                                    update_method, adam_mode, batch_size, window_size, num_random_entities,
                                    similarity_batch_size, float(regularization_lambda), float(text_entity_weight),
                                    float(similarity_weight), seed)
        assert self.h, "unknown objective"

    def __del__(self):
        try:
            self.L.ref2_destroy(self.h)
        except Exception:
            pass

    @property
    def rng_state(self):
        return self.L.ref2_get_rng_state(self.h)

    def _pf(self, a):
        return a.ctypes.data_as(ctypes.POINTER(self.L._ct))

    def fill_text(self, features, labels, feature_weights, weights):
        f = np.ascontiguousarray(features, dtype=np.int64).ravel()
        fw = np.ascontiguousarray(feature_weights, dtype=self.dtype).ravel()
        lab = np.ascontiguousarray(labels, dtype=np.int64).ravel()
        w = np.ascontiguousarray(weights, dtype=self.dtype).ravel()
        pl = ctypes.POINTER(ctypes.c_long)
        self.L.ref2_fill_text(self.h, f.ctypes.data_as(pl), self._pf(fw), lab.ctypes.data_as(pl), self._pf(w), lab.size)

    def fill_pairs(self, pairs, weights):
        ids = np.ascontiguousarray(pairs, dtype=np.int64).ravel()
        w = np.ascontiguousarray(weights, dtype=self.dtype).ravel()
        assert ids.size == 2 * w.size
        self.L.ref2_fill_pairs(self.h, ids.ctypes.data_as(ctypes.POINTER(ctypes.c_long)), self._pf(w), w.size)

    def get(self, name):
        n = self.L.ref2_tensor_size(self.h, name.encode())
        if n < 0:
            raise KeyError(name)
        out = np.zeros(n, dtype=self.dtype)
        assert self.L.ref2_get_tensor(self.h, name.encode(), self._pf(out)) == 0
        return out

    def forward(self):
        self.L.ref2_forward(self.h)

    def get_cost(self):
        return self.L.ref2_get_cost(self.h)

    def scaled_lambda(self):
        return self.L.ref2_scaled_regularization_lambda(self.h)

    def compute_gradients(self):
        self.L.ref2_compute_gradients(self.h)

    def update(self, lr, scaled_lambda):
        self.L.ref2_update(self.h, lr, scaled_lambda)
