"""Host-side mirror of the reference's plugin surface for the TextEntity path.

Names follow the reference (include/cuNVSM/model.h, data.h, proto/nvsm.proto):
``ModelDesc``, ``TrainConfig``, ``Batch`` and ``Model`` with ``initialize /
compute_cost / compute_gradients / update / backprop / get_cost / infer / get_data``.
Everything here is argument marshalling over the C ABI in include/nvsm_b200.h — all
arithmetic runs in the hand-written CUDA kernels of libnvsm_b200.so.
"""
import ctypes
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import NvsmConfig, NvsmError, check  # noqa: F401

TANH, HARD_TANH = 0, 1
SGD, ADAGRAD, ADAM = 0, 1, 2
SPARSE, DENSE_UPDATE, DENSE_UPDATE_DENSE_VARIANCE = 1, 2, 3
GEMM_FP32, GEMM_TF32, GEMM_3XTF32 = 0, 1, 2

# cuNVSMTrainModel --update_method values (cpp/main.cu:479-485)
UPDATE_METHODS = {
    "sgd": (SGD, 0),
    "adagrad": (ADAGRAD, 0),
    "sparse_adam": (ADAM, SPARSE),
    "dense_adam": (ADAM, DENSE_UPDATE),
    "full_adam": (ADAM, DENSE_UPDATE_DENSE_VARIANCE),
}
NONLINEARITIES = {"tanh": TANH, "hard_tanh": HARD_TANH}

SPARSE_LOCAL, SPARSE_ALLGATHER = 0, 1
# Model<...::Objective> instantiations of the reference (cpp/model.cu:222-228)
TEXT_ENTITY, ENTITY_ENTITY, TERM_TERM, TEXT_ENTITY_ENTITY_ENTITY, TEXT_ENTITY_TERM_TERM = 0, 1, 2, 3, 4

WORD_REPRS = "word_representations-representations"
ENTITY_REPRS = "entity_representations-representations"
TRANSFORM = "word_entity_mapping-transform"
BIAS = "word_entity_mapping-bias"


@dataclass
class ModelDesc:
    """lse::ModelDesc (proto/nvsm.proto:7-29)."""
    word_repr_size: int = 4
    entity_repr_size: int = 4
    batch_normalization: bool = False
    nonlinearity: int = TANH
    clip_sigmoid: bool = False
    bias_negative_samples: bool = False
    l2_normalize_phrase_reprs: bool = False
    l2_normalize_entity_reprs: bool = False


@dataclass
class TrainConfig:
    """lse::TrainConfig (proto/nvsm.proto:31-71)."""
    batch_size: int = 1024
    window_size: int = 8
    num_random_entities: int = 1
    regularization_lambda: float = 0.01
    learning_rate: float = 0.0
    update_method: int = SGD
    adam_mode: int = SPARSE
    text_entity_weight: float = 1.0
    entity_entity_weight: float = 0.0
    term_term_weight: float = 0.0
    num_epochs: int = 1


def _pl(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_long))


def _pf(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _fw(batch):
    """feature_weights / weights of a Batch for the C ABI: NULL when the batch was filled with uniform weights."""
    return (None if getattr(batch, "uniform_feature_weights_", False) else _pf(batch.feature_weights_),
            None if getattr(batch, "uniform_weights_", False) else _pf(batch.weights_))


def _pinned(shape, dtype):
    """Page-locked host array (falls back to pageable memory when torch/CUDA is absent)."""
    try:
        import torch
        if torch.cuda.is_available():
            t = torch.empty(shape, dtype={np.int64: torch.int64, np.float32: torch.float32}[dtype]).pin_memory()
            return t.numpy(), t
    except Exception:
        pass
    return np.zeros(shape, dtype=dtype), None


class Batch:
    """TextEntity::Batch (include/cuNVSM/data.h:114-177): host memory of one step.

    features_[B*n] int64 (n word ids per n-gram), feature_weights_[B*n] float32,
    labels_[B] int64 (positive document), weights_[B] float32.
    """

    def __init__(self, batch_size, window_size, pinned=True):
        self.batch_size_, self.window_size_ = int(batch_size), int(window_size)
        self._keep = []
        def alloc(n, dt):
            if pinned:
                a, owner = _pinned((n,), dt)
                self._keep.append(owner)
                return a
            return np.zeros(n, dtype=dt)
        self.features_ = alloc(self.batch_size_ * self.window_size_, np.int64)
        self.feature_weights_ = alloc(self.batch_size_ * self.window_size_, np.float32)
        self.labels_ = alloc(self.batch_size_, np.int64)
        self.weights_ = alloc(self.batch_size_, np.float32)
        self.num_instances_ = 0
        self.uniform_feature_weights_ = self.uniform_weights_ = False

    def window_size(self):
        return self.window_size_

    def num_instances(self):
        return self.num_instances_

    def maximum_size(self):
        return self.batch_size_

    def full(self):
        return self.num_instances_ == self.batch_size_

    def empty(self):
        return self.num_instances_ == 0

    def clear(self):
        self.num_instances_ = 0

    def push_instance(self, features, feature_weights, object_id, weight):
        """DataSource::push_instance (cpp/data.cu:94-124)."""
        assert not self.full() and len(features) == self.window_size_
        i, n = self.num_instances_, self.window_size_
        self.features_[i * n:(i + 1) * n] = features
        self.feature_weights_[i * n:(i + 1) * n] = 1.0 if feature_weights is None else feature_weights
        self.labels_[i] = object_id
        self.weights_[i] = weight
        self.num_instances_ += 1
        self.uniform_feature_weights_ = self.uniform_weights_ = False

    def fill(self, features, labels, feature_weights=None, weights=None):
        features = np.asarray(features, dtype=np.int64).reshape(-1, self.window_size_)
        B = features.shape[0]
        assert B <= self.batch_size_
        self.features_[:B * self.window_size_] = features.ravel()
        self.feature_weights_[:B * self.window_size_] = 1.0 if feature_weights is None else np.asarray(feature_weights, dtype=np.float32).ravel()
        self.labels_[:B] = labels
        self.weights_[:B] = 1.0 if weights is None else weights
        self.num_instances_ = B
        # uniform weighting (the reference's default): the arrays hold 1.0 and need not cross PCIe (NULL over the C ABI)
        self.uniform_feature_weights_, self.uniform_weights_ = feature_weights is None, weights is None
        return self


class RNG:
    """std::minstd_rand0 state holder (include/cuNVSM/base.h:36); the engine itself runs in C++."""

    def __init__(self, seed=1):
        self.seed(seed)

    def seed(self, s):
        s = int(s) % 2147483647
        self.state = s if s != 0 else 1


class ForwardResult:
    """Handle on the forward state living in the model's device workspace."""

    def __init__(self, model, num_instances):
        self._model, self.batch_size_ = model, num_instances

    def get_cost(self):
        out = ctypes.c_float()
        check(self._model.L.nvsm_get_cost(self._model.h, ctypes.byref(out)))
        return out.value

    def scaled_regularization_lambda(self):
        return self._model.L.nvsm_scaled_regularization_lambda(self._model.h)

    def get_similarity_probs(self):
        return self._model.get_tensor("similarity_probs")


class SimilarityBatch:
    """RepresentationSimilarity::Batch (include/cuNVSM/data.h:560-614): features_ [2*batch] pair ids, weights_ [batch]."""

    def __init__(self, batch_size):
        self.batch_size_ = int(batch_size)
        self.features_, keep_f = _pinned((2 * self.batch_size_,), np.int64)
        self.weights_, keep_w = _pinned((self.batch_size_,), np.float32)
        self._keep = [keep_f, keep_w]
        self.num_instances_ = 0

    def fill(self, pairs, weights=None):
        pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
        N = pairs.shape[0]
        assert N <= self.batch_size_
        self.features_[:2 * N] = pairs.ravel()
        self.weights_[:N] = 1.0 if weights is None else np.asarray(weights, dtype=np.float32)
        self.num_instances_ = N
        return self

    def num_instances(self):
        return self.num_instances_


class SimilarityForwardResult:
    """RepresentationSimilarity::ForwardResult: get_cost / scaled_regularization_lambda / get_similarity_probs."""

    def __init__(self, model, num_instances):
        self.model, self.num_instances = model, num_instances

    def get_cost(self):
        c = ctypes.c_float()
        check(self.model.L.nvsm_similarity_get_cost(self.model.h, ctypes.byref(c)))
        return float(c.value)

    def scaled_regularization_lambda(self):
        return float(self.model.L.nvsm_similarity_scaled_regularization_lambda(self.model.h))

    def get_similarity_probs(self):
        return self.model.get_tensor("similarity_pair_probs")


class MultiForwardResult:
    """MultiForwardResultBase (cpp/intermediate_results.cu:186-240): plain averages over the constituents."""

    def __init__(self, *results):
        self.results = results

    def get_cost(self):
        return float(np.float32(sum(np.float32(r.get_cost()) for r in self.results)) / np.float32(len(self.results)))

    def scaled_regularization_lambda(self):
        return float(np.float32(sum(np.float32(r.scaled_regularization_lambda()) for r in self.results))
                     / np.float32(len(self.results)))


class Gradients:
    def __init__(self, model):
        self._model = model

    def get(self, name):
        return self._model.get_tensor(name)


class Model:
    """Model<TextEntity::Objective> (include/cuNVSM/model.h:76-130) over the C ABI."""

    def __init__(self, num_words, num_entities, desc, train_config, device=0, gemm_mode=GEMM_FP32,
                 num_batch_slots=1, max_batch_size=None, objective=TEXT_ENTITY, max_similarity_batch_size=None):
        self.L = _lib.load()
        self.desc, self.train_config = desc, train_config
        self.num_words, self.num_entities = int(num_words), int(num_entities)
        cfg = NvsmConfig()
        cfg.num_words, cfg.num_entities = self.num_words, self.num_entities
        cfg.word_repr_size, cfg.entity_repr_size = desc.word_repr_size, desc.entity_repr_size
        cfg.nonlinearity = desc.nonlinearity
        cfg.batch_normalization = int(desc.batch_normalization)
        cfg.clip_sigmoid = int(desc.clip_sigmoid)
        cfg.bias_negative_samples = int(desc.bias_negative_samples)
        cfg.l2_normalize_phrase_reprs = int(desc.l2_normalize_phrase_reprs)
        cfg.l2_normalize_entity_reprs = int(desc.l2_normalize_entity_reprs)
        cfg.update_method, cfg.adam_mode = train_config.update_method, train_config.adam_mode
        cfg.num_random_entities = train_config.num_random_entities
        cfg.max_batch_size = int(max_batch_size or train_config.batch_size)
        cfg.window_size = train_config.window_size
        cfg.regularization_lambda = train_config.regularization_lambda
        cfg.device, cfg.gemm_mode, cfg.num_batch_slots = device, gemm_mode, num_batch_slots
        self.objective = cfg.objective = objective
        cfg.text_entity_weight = train_config.text_entity_weight
        cfg.similarity_weight = (train_config.entity_entity_weight if objective in (ENTITY_ENTITY, TEXT_ENTITY_ENTITY_ENTITY)
                                 else train_config.term_term_weight)
        cfg.max_similarity_batch_size = int(max_similarity_batch_size or train_config.batch_size)
        self.cfg = cfg
        self.h = ctypes.c_void_p()
        check(self.L.nvsm_create(ctypes.byref(cfg), ctypes.byref(self.h)))
        self._keepalive = None

    def close(self):
        if getattr(self, "h", None):
            self.L.nvsm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- parameters -----------------------------------------------------------------
    def initialize(self, rng):
        st = ctypes.c_ulong(rng.state)
        check(self.L.nvsm_initialize(self.h, ctypes.byref(st)))
        rng.state = st.value

    def tensor_size(self, name):
        return self.L.nvsm_tensor_size(self.h, name.encode())

    def get_tensor(self, name):
        n = self.tensor_size(name)
        if n < 0:
            raise KeyError(name)
        out = np.zeros(n, dtype=np.float32)
        check(self.L.nvsm_get_tensor(self.h, name.encode(), _pf(out), n))
        return out

    def set_tensor(self, name, value):
        value = np.ascontiguousarray(value, dtype=np.float32).ravel()
        check(self.L.nvsm_set_tensor(self.h, name.encode(), _pf(value), value.size))

    def get_data(self):
        """ModelBase::get_data (cpp/model.cu:64-93): name -> [objects, dim] array."""
        d = self.desc
        return {
            WORD_REPRS: self.get_tensor(WORD_REPRS).reshape(self.num_words, d.word_repr_size),
            ENTITY_REPRS: self.get_tensor(ENTITY_REPRS).reshape(self.num_entities, d.entity_repr_size),
            TRANSFORM: self.get_tensor(TRANSFORM).reshape(d.word_repr_size, d.entity_repr_size),
            BIAS: self.get_tensor(BIAS).reshape(1, d.entity_repr_size),
        }

    def num_parameters(self):
        d = self.desc
        return (self.num_words * d.word_repr_size + self.num_entities * d.entity_repr_size +
                d.word_repr_size * d.entity_repr_size + d.entity_repr_size)

    def increment_parameter(self, name, idx, epsilon):
        check(self.L.nvsm_increment_parameter(self.h, name.encode(), idx, epsilon))

    # --- the step ----------------------------------------------------------------------
    def generate_labels(self, labels, rng):
        """Objective::generate_labels (cpp/objective.cu:5-28): bit-exact host sampler."""
        labels = np.ascontiguousarray(labels, dtype=np.int64)
        z = self.train_config.num_random_entities
        out = np.zeros(labels.size * (z + 1), dtype=np.int64)
        st = ctypes.c_ulong(rng.state)
        cdf = getattr(self, "_cdf", None)
        if cdf is not None:
            check(self.L.nvsm_generate_labels_cdf(_pl(labels), labels.size, z, cdf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                                  cdf.size, ctypes.byref(st), _pl(out)))
        else:
            check(self.L.nvsm_generate_labels(_pl(labels), labels.size, z, self.num_entities, ctypes.byref(st), _pl(out)))
        rng.state = st.value
        return out

    def set_negative_distribution(self, cdf):
        """Plug a skewed negative sampler in where the reference has LabelGenerator (include/cuNVSM/labels.h:7-18):
        inverse-CDF draws over ``cdf`` [num_entities] (see zipf_cdf) on the same shared engine, one engine output per
        negative — host (generate_labels) and device (step_sampled) draw identical ids. ``None`` restores the
        reference's UniformLabelGenerator."""
        if cdf is None:
            self._cdf = None
            check(self.L.nvsm_sampler_set_cdf(self.h, None, 0))
            return
        cdf = np.ascontiguousarray(cdf, dtype=np.float64)
        check(self.L.nvsm_sampler_set_cdf(self.h, cdf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), cdf.size))
        self._cdf = cdf

    def compute_cost(self, batch, rng=None, entity_ids=None):
        """Model::compute_cost (cpp/objective.cu:30-313). Negatives are drawn from ``rng``
        exactly like the reference unless ``entity_ids`` [B*(z+1)] is supplied."""
        B, n = batch.num_instances_, batch.window_size_
        assert n == self.train_config.window_size
        if entity_ids is None:
            entity_ids = self.generate_labels(batch.labels_[:B], rng)
        entity_ids = np.ascontiguousarray(entity_ids, dtype=np.int64)
        assert entity_ids.size == B * (self.train_config.num_random_entities + 1)
        self._keepalive = (batch, entity_ids)
        fw, w = _fw(batch)
        check(self.L.nvsm_compute_cost(self.h, _pl(batch.features_), fw, _pl(entity_ids), w, B))
        return ForwardResult(self, B)

    def compute_gradients(self, result=None):
        check(self.L.nvsm_compute_gradients(self.h))
        return Gradients(self)

    # --- RepresentationSimilarity objective (EntityEntity / TermTerm and the mixtures) ----------
    def similarity_compute_cost(self, batch):
        """RepresentationSimilarity::Objective::compute_cost (cpp/objective.cu:487-573) on a SimilarityBatch."""
        N = batch.num_instances_
        self._keepalive_pairs = batch
        check(self.L.nvsm_similarity_compute_cost(self.h, _pl(batch.features_), _pf(batch.weights_), N))
        return SimilarityForwardResult(self, N)

    def compute_cost_mixture(self, text_batch, similarity_batch, rng=None, entity_ids=None):
        """TextEntity{EntityEntity,TermTerm}::Objective::compute_cost (cpp/objective.cu:713-724,762-773): both
        constituents on their own batch; the result averages their costs and scaled lambdas."""
        text = self.compute_cost(text_batch, rng, entity_ids)
        sim = self.similarity_compute_cost(similarity_batch)
        return MultiForwardResult(text, sim)

    def update(self, gradients, learning_rate, scaled_regularization_lambda):
        check(self.L.nvsm_update(self.h, learning_rate, scaled_regularization_lambda))

    def backprop(self, result, learning_rate):
        g = self.compute_gradients(result)
        self.update(g, learning_rate, result.scaled_regularization_lambda())

    def get_cost(self, batch, rng_state, rng):
        """Model::get_cost (cpp/model.cu:154-174): optionally restore the RNG state first."""
        if rng_state is not None:
            rng.state = rng_state
        return self.compute_cost(batch, rng).get_cost()

    def train_step(self, batch, entity_ids, learning_rate):
        """compute_cost + compute_gradients + update on host buffers, no synchronisation."""
        B = batch.num_instances_
        self._keepalive = (batch, entity_ids)
        fw, w = _fw(batch)
        check(self.L.nvsm_train_step(self.h, _pl(batch.features_), fw, _pl(entity_ids), w, B, learning_rate))

    # --- device sampler ---------------------------------------------------------------------
    def sampler_seed(self, rng):
        """Move the engine state to the device (nvsm_sampler_seed)."""
        check(self.L.nvsm_sampler_seed(self.h, rng.state))

    def sampler_state(self):
        st = ctypes.c_ulong()
        check(self.L.nvsm_sampler_state(self.h, ctypes.byref(st)))
        return st.value

    def step_sampled(self, batch, learning_rate, train=True):
        """Upload the batch, draw the negatives on the device (bit-exact with the host sampler) and run
        compute_cost (+ compute_gradients + update when train). No synchronisation."""
        self._keepalive = (batch,)
        fw, w = _fw(batch)
        check(self.L.nvsm_step_sampled(self.h, _pl(batch.features_), fw, _pl(batch.labels_), w, batch.num_instances_,
                                       learning_rate, int(train)))

    def generate_labels_device(self, labels, rng, z=None, num_objects=None):
        """nvsm_generate_labels on the device (bit-exact): returns ids [B*(z+1)], advances rng."""
        labels = np.ascontiguousarray(labels, dtype=np.int64)
        z = self.train_config.num_random_entities if z is None else z
        D = self.num_entities if num_objects is None else num_objects
        out = np.zeros(labels.size * (z + 1), dtype=np.int64)
        st = ctypes.c_ulong(rng.state)
        check(self.L.nvsm_generate_labels_device(self.h, _pl(labels), labels.size, z, D, ctypes.byref(st), _pl(out)))
        rng.state = st.value
        return out

    def entity_ids(self, num_instances):
        n = num_instances * (self.train_config.num_random_entities + 1)
        out = np.zeros(n, dtype=np.int64)
        check(self.L.nvsm_get_entity_ids(self.h, _pl(out), n))
        return out

    def stage_batch(self, slot, batch, entity_ids):
        entity_ids = np.ascontiguousarray(entity_ids, dtype=np.int64)
        fw, w = _fw(batch)
        check(self.L.nvsm_stage_batch(self.h, slot, _pl(batch.features_), fw, _pl(entity_ids), w, batch.num_instances_))

    def compute_cost_staged(self, slot):
        check(self.L.nvsm_compute_cost_staged(self.h, slot))
        return ForwardResult(self, None)

    def train_step_staged(self, slot, learning_rate):
        check(self.L.nvsm_train_step_staged(self.h, slot, learning_rate))

    def last_cost(self, steps_back=0):
        out = ctypes.c_float()
        check(self.L.nvsm_read_cost(self.h, steps_back, ctypes.byref(out)))
        return out.value

    def infer(self, words, window_size):
        """Model::infer (cpp/model.cu:105-133): words [N][window] -> [N, d_d]."""
        words = np.ascontiguousarray(words, dtype=np.int64).reshape(-1, window_size)
        out = np.zeros((words.shape[0], self.desc.entity_repr_size), dtype=np.float32)
        check(self.L.nvsm_infer(self.h, _pl(words), words.shape[0], window_size, _pf(out)))
        return out

    # --- plumbing ----------------------------------------------------------------------
    def set_stream(self, cuda_stream_handle):
        check(self.L.nvsm_set_stream(self.h, ctypes.c_void_p(cuda_stream_handle)))

    def synchronize(self):
        check(self.L.nvsm_synchronize(self.h))

    def set_profiling(self, on):
        check(self.L.nvsm_set_profiling(self.h, int(on)))

    def timeline(self, capacity=4096):
        """Intervals (phase name, start ms, end ms) since set_profiling(2), overlaps between streams kept."""
        ph = (ctypes.c_int * capacity)()
        a = (ctypes.c_float * capacity)()
        b = (ctypes.c_float * capacity)()
        n = self.L.nvsm_get_timeline(self.h, ph, a, b, capacity)
        if n < 0:
            check(1)
        return [(self.L.nvsm_phase_name(ph[i]).decode(), a[i], b[i]) for i in range(min(n, capacity))]

    def reset_phase_ms(self):
        check(self.L.nvsm_reset_phase_ms(self.h))

    def phase_ms(self):
        n = self.L.nvsm_num_phases()
        out = (ctypes.c_float * n)()
        check(self.L.nvsm_get_phase_ms(self.h, out, n))
        return {self.L.nvsm_phase_name(i).decode(): out[i] for i in range(n)}

    def bench_memory(self, kind, table_bytes, row_floats=256, rows_per_item=10, items=51200, iters=20):
        """Roofline denominators measured in place (nvsm_bench_memory): kind 0 = GB/s of a plain row gather out of a
        `table_bytes` table (L2-resident when it fits), kind 1 = GB/s of a streaming copy."""
        out = ctypes.c_float()
        check(self.L.nvsm_bench_memory(self.h, kind, int(table_bytes), row_floats, rows_per_item, int(items), iters, ctypes.byref(out)))
        return out.value

    def kernel_launches(self):
        return self.L.nvsm_kernel_launches(self.h)

    def comm_init(self, unique_id, num_ranks, rank):
        check(self.L.nvsm_comm_init(self.h, unique_id, num_ranks, rank))

    def comm_peer_export(self):
        buf = ctypes.create_string_buffer(128)
        check(self.L.nvsm_comm_peer_export(self.h, buf))
        return buf.raw

    def comm_peer_import(self, blobs):
        """blobs: the 128-byte exports of all ranks, in rank order."""
        data = b"".join(blobs)
        check(self.L.nvsm_comm_peer_import(self.h, data))

    def peer_ready_override_off(self):
        check(self.L.nvsm_comm_peer_disable(self.h))

    def comm_peer_status(self):
        ready, err = ctypes.c_int(), ctypes.c_int()
        check(self.L.nvsm_comm_peer_status(self.h, ctypes.byref(ready), ctypes.byref(err)))
        return bool(ready.value), int(err.value)

    def comm_set_sparse_mode(self, mode):
        """SPARSE_LOCAL: per-rank local table updates; SPARSE_ALLGATHER: every replica applies the updates of
        the whole global batch (the single-GPU trajectory)."""
        check(self.L.nvsm_comm_set_sparse_mode(self.h, mode))


def zipf_cdf(num_objects, exponent=1.0):
    """Cumulative Zipf(s) distribution over ids 0..num_objects-1 (id k has weight (k+1)^-s), ending at exactly 1.0 —
    the argument of Model.set_negative_distribution / nvsm_sampler_set_cdf."""
    w = 1.0 / np.power(np.arange(1, num_objects + 1, dtype=np.float64), float(exponent))
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    cdf[-1] = 1.0
    return cdf


def comm_unique_id():
    buf = ctypes.create_string_buffer(128)
    check(_lib.load().nvsm_comm_unique_id(buf))
    return buf.raw
