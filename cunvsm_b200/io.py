"""Reading what cuNVSMTrainModel writes: `<output>_meta` (lse.Metadata, proto/nvsm.proto:91-108) and the model dump
`<output>_<epoch>.<dataset>.npy` (the four datasets of the reference's HDF5 file, cpp/hdf5.cu:26-53, row-major
[objects, dim]). `DumpedModel` exposes them under the attribute names of the reference's Python model class
(py/nvsm/base.py:165-236: word_representations, object_representations, transform_matrix, transform_bias, the id
mappings) with its query-side arithmetic (query_representation, infer), so downstream code written against that class
keeps working without HDF5 / protobuf installed. Host-side only; no GPU, no library call.
"""
import numpy as np

DATASETS = ("word_representations-representations", "entity_representations-representations",
            "word_entity_mapping-transform", "word_entity_mapping-bias")


def _varint(buf, pos):
    value, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise ValueError("truncated varint")
        b = buf[pos]
        pos += 1
        value |= (b & 0x7F) << shift
        if not b & 0x80:
            return value, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def _int32(value):
    """int32 fields travel sign-extended to 64 bits."""
    value &= (1 << 64) - 1
    return value - (1 << 64) if value >= 1 << 63 else value


def _fields(buf):
    """(field number, wire type, value) of one message; value = int for varints, bytes for length-delimited."""
    pos = 0
    while pos < len(buf):
        key, pos = _varint(buf, pos)
        field, wire = key >> 3, key & 7
        if wire == 0:
            value, pos = _varint(buf, pos)
        elif wire == 2:
            size, pos = _varint(buf, pos)
            if size > len(buf) - pos:
                raise ValueError("truncated field")
            value, pos = bytes(buf[pos:pos + size]), pos + size
        elif wire == 1:
            value, pos = bytes(buf[pos:pos + 8]), pos + 8
        elif wire == 5:
            value, pos = bytes(buf[pos:pos + 4]), pos + 4
        else:
            raise ValueError("unsupported wire type %d" % wire)
        if pos > len(buf):
            raise ValueError("truncated field")
        yield field, wire, value


def parse_metadata(data):
    """lse.Metadata -> {"term": [(index_term_id, model_term_id, term_frequency)], "object": [(index_object_id,
    model_object_id)], "total_terms": int} (unknown fields are skipped, absent scalars are 0: proto3)."""
    meta = {"term": [], "object": [], "total_terms": 0}
    for field, wire, value in _fields(memoryview(data)):
        if field == 1 and wire == 2:
            v = {1: 0, 2: 0, 3: 0}
            for f, w, x in _fields(memoryview(value)):
                if w == 0 and f in v:
                    v[f] = _int32(x)
            meta["term"].append((v[1], v[2], v[3]))
        elif field == 2 and wire == 2:
            v = {1: 0, 2: 0}
            for f, w, x in _fields(memoryview(value)):
                if w == 0 and f in v:
                    v[f] = _int32(x)
            meta["object"].append((v[1], v[2]))
        elif field == 3 and wire == 0:
            meta["total_terms"] = _int32(value)
    return meta


def load_meta(path):
    """py/nvsm/base.py:load_meta: reads `<path>_meta`."""
    with open("%s_meta" % path, "rb") as f:
        return parse_metadata(f.read())


class DumpedModel:
    """The reference's `NVSM` model object (py/nvsm/base.py:165-330) over a .npy dump.

    ``bias_coefficient`` defaults to 0.0 like the reference (py/nvsm/base.py:171), where the projection bias then
    contributes nothing to the query projection (its `if not bias_coefficient != 0.0` keeps a zero vector for 0.0 and
    drops the bias otherwise). Here a non-zero coefficient scales and ADDS the trained bias, which is what
    Model::infer (cpp/model.cu:105-133) computes with coefficient 1.0."""

    def __init__(self, meta, tensors, bias_coefficient=0.0, nonlinearity=np.tanh, self_information=False, strict=False):
        self.total_terms = meta["total_terms"]
        self.self_information, self.nonlinearity, self.strict = self_information, nonlinearity, strict
        self.word_representations = tensors[DATASETS[0]]
        self.num_terms, self.term_repr_size = self.word_representations.shape
        self.term_mapping, self.inv_term_mapping, self.inv_term_id_to_term_freq = {}, {}, {}
        for index_term_id, model_term_id, term_frequency in meta["term"]:
            assert index_term_id not in self.term_mapping and model_term_id < self.num_terms
            self.term_mapping[index_term_id] = model_term_id
            self.inv_term_mapping[model_term_id] = index_term_id
            self.inv_term_id_to_term_freq[model_term_id] = term_frequency
        self.object_representations = tensors[DATASETS[1]]
        self.num_objects, self.object_repr_size = self.object_representations.shape
        self.object_mapping, self.inv_object_mapping = {}, {}
        for index_object_id, model_object_id in meta["object"]:
            assert model_object_id not in self.object_mapping and model_object_id < self.num_objects
            self.object_mapping[model_object_id] = index_object_id
            self.inv_object_mapping[index_object_id] = model_object_id
        self.transform_matrix = tensors[DATASETS[2]]
        assert self.transform_matrix.shape == (self.term_repr_size, self.object_repr_size)
        self.transform_bias = bias_coefficient * tensors[DATASETS[3]].ravel() if bias_coefficient != 0.0 else None

    def query_representation(self, index_term_ids):
        """Average (optionally self-information weighted) word representation of the in-vocabulary query terms."""
        terms = [self.term_mapping[t] for t in index_term_ids if t in self.term_mapping]
        if not terms or (self.strict and len(terms) < len(index_term_ids)):
            return None
        weights = None
        if self.self_information:
            weights = [-np.log(self.inv_term_id_to_term_freq[t] / self.total_terms) for t in terms]
        return np.average(self.word_representations[terms, :], axis=0, weights=weights)

    def infer(self, query_repr):
        """Projection into the document space (Model::infer, cpp/model.cu:105-133: no batch-norm at inference)."""
        if query_repr is None:
            return None
        projected = np.dot(query_repr, self.transform_matrix)
        if self.transform_bias is not None:
            projected = projected + self.transform_bias
        return self.nonlinearity(projected) if self.nonlinearity is not None else projected

    def rank(self, index_term_ids, results_requested=10):
        """(cosine similarity, index object id) of the closest documents, best first."""
        q = self.infer(self.query_representation(index_term_ids))
        if q is None:
            return None
        docs = self.object_representations
        sims = docs @ q / (np.linalg.norm(docs, axis=1) * np.linalg.norm(q) + 1e-30)
        order = np.argsort(-sims)[:results_requested]
        return [(float(sims[i]), self.object_mapping.get(int(i), int(i))) for i in order]


def load_model(meta, path, epoch, **kwargs):
    """py/nvsm/base.py:load_model, over `<path>_<epoch>.<dataset>.npy`."""
    return DumpedModel(meta, {name: np.load("%s_%s.%s.npy" % (path, epoch, name)) for name in DATASETS}, **kwargs)
