"""`_meta` files (lse.Metadata, reference: proto/nvsm.proto:91-108, written by cpp/main.cu:527-537 and read by
py/nvsm/base.py:load_meta): the hand-written wire-format encoder / parser of include/cuNVSM/nvsm.pb.h against the
protobuf runtime, both directions. CPU only."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "cpp", "cuNVSMMeta")


def metadata_class():
    """lse.Metadata built from a descriptor that restates proto/nvsm.proto:91-108 (protoc is not in the image)."""
    pytest.importorskip("google.protobuf")
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

    fd = descriptor_pb2.FileDescriptorProto(name="nvsm_meta_test.proto", package="lse", syntax="proto3")
    meta = fd.message_type.add(name="Metadata")
    I32, MSG = descriptor_pb2.FieldDescriptorProto.TYPE_INT32, descriptor_pb2.FieldDescriptorProto.TYPE_MESSAGE
    OPT, REP = descriptor_pb2.FieldDescriptorProto.LABEL_OPTIONAL, descriptor_pb2.FieldDescriptorProto.LABEL_REPEATED
    term = meta.nested_type.add(name="TermInfo")
    for i, n in enumerate(("index_term_id", "model_term_id", "term_frequency"), 1):
        term.field.add(name=n, number=i, type=I32, label=OPT)
    obj = meta.nested_type.add(name="ObjectInfo")
    for i, n in enumerate(("index_object_id", "model_object_id"), 1):
        obj.field.add(name=n, number=i, type=I32, label=OPT)
    meta.field.add(name="term", number=1, type=MSG, label=REP, type_name=".lse.Metadata.TermInfo")
    meta.field.add(name="object", number=2, type=MSG, label=REP, type_name=".lse.Metadata.ObjectInfo")
    meta.field.add(name="total_terms", number=3, type=I32, label=OPT)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("lse.Metadata"))


def tool():
    if not os.path.exists(TOOL):
        subprocess.run(["make", "-C", os.path.join(ROOT, "cpp"), "cuNVSMMeta"], check=True, capture_output=True)
    return TOOL


def test_written_meta_parses_with_protobuf_runtime(tmp_path):
    rng = np.random.default_rng(3)
    n, N, V, D = 3, 700, 300, 41          # V > 127: multi-byte varints
    words = rng.integers(0, V, size=(N, n))
    words[0, 0] = V - 1
    docs = rng.integers(0, D, size=N)
    docs[1] = D - 1
    path = tmp_path / "ngrams.txt"
    with open(path, "w") as f:
        f.write("# comment\n")
        for i in range(N):
            f.write("%d %s\n" % (docs[i], " ".join(map(str, words[i]))))
    out = subprocess.run([tool(), "write", str(path), str(n), str(tmp_path / "model")], check=True, capture_output=True,
                         text=True).stdout
    assert out.split() == ["terms", str(V), "objects", str(D), "total_terms", str(N * n)]
    meta = metadata_class()()
    meta.ParseFromString((tmp_path / "model_meta").read_bytes())
    assert len(meta.term) == V and len(meta.object) == D and meta.total_terms == N * n
    freq = np.bincount(words.ravel(), minlength=V)
    for i, t in enumerate(meta.term):
        assert (t.index_term_id, t.model_term_id, t.term_frequency) == (i, i, freq[i])
    for j, o in enumerate(meta.object):
        assert (o.index_object_id, o.model_object_id) == (j, j)
    # byte-identical with the runtime's own (deterministic) serialisation
    assert meta.SerializeToString(deterministic=True) == (tmp_path / "model_meta").read_bytes()


def test_parser_reads_protobuf_runtime_output(tmp_path):
    """A Metadata file as the reference writes it (arbitrary index ids, negative ids, zero fields omitted)."""
    meta = metadata_class()()
    expect = []
    for index_id, model_id, tf in ((907, 0, 5), (12, 1, 0), (70000, 2, 123456), (-3, 3, 1)):
        meta.term.add(index_term_id=index_id, model_term_id=model_id, term_frequency=tf)
        expect.append("term %d %d %d" % (index_id, model_id, tf))
    for index_id, model_id in ((4000001, 0), (17, 1), (0, 2)):
        meta.object.add(index_object_id=index_id, model_object_id=model_id)
        expect.append("object %d %d" % (index_id, model_id))
    meta.total_terms = 2 ** 31 - 1
    expect.append("total_terms %d" % (2 ** 31 - 1))
    path = tmp_path / "ref_meta"
    path.write_bytes(meta.SerializeToString())
    out = subprocess.run([tool(), "print", str(path)], check=True, capture_output=True, text=True).stdout
    assert out.strip().split("\n") == expect


def test_parser_rejects_truncated_file(tmp_path):
    meta = metadata_class()()
    meta.term.add(index_term_id=300, model_term_id=1, term_frequency=9)
    data = meta.SerializeToString()
    path = tmp_path / "bad_meta"
    path.write_bytes(data[:-2])
    r = subprocess.run([tool(), "print", str(path)], capture_output=True, text=True)
    assert r.returncode != 0 and "not an lse.Metadata message" in (r.stderr + r.stdout)


def test_python_loader_reads_meta_and_dump(tmp_path):
    """cunvsm_b200/io.py: the `_meta` parser agrees with the protobuf runtime, and DumpedModel reproduces the
    reference's query-side arithmetic (py/nvsm/base.py: query_representation / infer) and the oracle's Model::infer."""
    from cunvsm_b200 import io
    from oracle import binding as O
    meta_cls = metadata_class()
    meta = meta_cls()
    V, D, dw, dd = 40, 12, 8, 6
    for k in range(V):
        meta.term.add(index_term_id=1000 + 3 * k, model_term_id=k, term_frequency=5 + k)
    for j in range(D):
        meta.object.add(index_object_id=70000 - j, model_object_id=j)
    meta.total_terms = sum(5 + k for k in range(V))
    out = str(tmp_path / "model")
    with open(out + "_meta", "wb") as f:
        f.write(meta.SerializeToString())
    parsed = io.load_meta(out)
    assert parsed["total_terms"] == meta.total_terms
    assert parsed["term"] == [(t.index_term_id, t.model_term_id, t.term_frequency) for t in meta.term]
    assert parsed["object"] == [(o.index_object_id, o.model_object_id) for o in meta.object]
    # negative ids and unknown fields survive
    assert io.parse_metadata(b"\x0a\x0b\x08\xfd\xff\xff\xff\xff\xff\xff\xff\xff\x01\x20\x07")["term"] == [(-3, 0, 0)]

    om = O.Model(V, D, dw, dd, nonlinearity=O.TANH, batch_normalization=False, dtype=np.float32)
    om.initialize(3)
    rng = np.random.default_rng(0)
    om.set("b", rng.normal(size=dd).astype(np.float32))
    names = dict(zip(io.DATASETS, ("W", "E", "T", "b")))
    for name, short in names.items():
        a = np.asarray(om.get(short), dtype=np.float32)
        shape = {"W": (V, dw), "E": (D, dd), "T": (dw, dd), "b": (1, dd)}[short]
        np.save("%s_%d.%s.npy" % (out, 2, name), a.reshape(shape))
    assert io.load_model(parsed, out, 2).transform_bias is None      # reference default: the bias plays no part
    model = io.load_model(parsed, out, 2, bias_coefficient=1.0)        # Model::infer adds it
    assert (model.num_terms, model.term_repr_size, model.num_objects, model.object_repr_size) == (V, dw, D, dd)
    query = [1000 + 3 * 4, 1000 + 3 * 9, 999999, 1000 + 3 * 17]          # one out-of-vocabulary term
    projected = model.infer(model.query_representation(query))
    expect = om.infer(np.array([[4, 9, 17]]), 3)[0]                          # gather-mean -> T p + b -> tanh
    np.testing.assert_allclose(projected, expect, rtol=2e-5, atol=1e-6)
    ranked = model.rank(query, 3)
    assert len(ranked) == 3 and ranked[0][0] >= ranked[1][0] >= ranked[2][0] and all(70000 - D < d <= 70000 for _, d in ranked)
    assert io.load_model(parsed, out, 2, strict=True).query_representation(query) is None
