"""-m gpu parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs, plus the reference's golden vectors and size-independent properties at the
BASELINE.json sizes.

Tolerances (fp32 arithmetic, NVSM_GEMM_FP32):
  * sampled / passed indices: bit-exact.
  * forward tensors, loss, gradients vs the float32 oracle: rtol 2e-4 with an absolute floor
    of 1e-6 x max|expected| (summation order differs: warp-shuffle trees, split-K, atomics).
  * parameters after optimiser steps: rtol 5e-4, same floor.
"""
import json
import os

import numpy as np
import pytest

import cunvsm_b200 as nv
from oracle import binding as O
from tests.util import assert_close, make_batch, twin_models

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_goldens.json")))
RTOL = 2e-4


def run_forward_backward(gm, om, rng, B, n, V, D, z, weighted=True, seed=3):
    nrng = np.random.default_rng(seed)
    f, fw, labels, w = make_batch(nrng, B, n, V, D, z, weighted)
    batch = nv.Batch(B, n).fill(f, labels, fw, w)
    state0 = rng.state
    ids_ref, _ = O.generate_labels(labels, z, D, state0)
    res = gm.compute_cost(batch, rng)
    ids = gm._keepalive[1]
    assert (ids == ids_ref).all(), "sampled entity ids must be bit-exact"
    cost = res.get_cost()
    ocost = om.compute_cost(f, fw, ids_ref, w, n)
    gm.compute_gradients(res)
    om.compute_gradients()
    return batch, ids_ref, cost, ocost


def compare_step_tensors(gm, om, cost, ocost, rtol=RTOL):
    assert abs(cost - ocost) <= rtol * abs(ocost) + 1e-7
    assert_close(gm.get_tensor("phrase_reprs"), om.get("P"), rtol, what="P")
    assert_close(gm.get_tensor("word_projections"), om.get("Y"), rtol, 2e-6, what="Y")
    assert_close(gm.get_tensor("similarity_probs"), om.get("probs"), rtol, what="probs")
    assert_close(gm.get_tensor("instance_multipliers"), om.get("mult"), rtol, 1e-5, what="mult")
    assert_close(gm.get_tensor("grad_transform"), om.get("gT"), rtol, 1e-5, what="gT")
    assert_close(gm.get_tensor("grad_bias"), om.get("gb"), rtol, 1e-5, what="gb")
    assert_close(gm.get_tensor("grad_phrase_reprs"), om.get("gP"), rtol, 1e-5, what="gP")
    assert_close(gm.get_tensor("grad_entity_repr"), om.get("gE"), rtol, 1e-5, what="gE")


def test_reference_golden_transform_backward():
    """cpp/model_tests.cu:341-466 through the CUDA path (fp32 => 1e-5 relative)."""
    g = GOLD["transform_backward"]; c = g["config"]
    desc = nv.ModelDesc(word_repr_size=c["word_repr_size"], entity_repr_size=c["entity_repr_size"],
                        bias_negative_samples=True)
    tc = nv.TrainConfig(batch_size=c["batch_size"], window_size=c["window_size"],
                        num_random_entities=c["num_random_entities"], regularization_lambda=0.01)
    m = nv.Model(c["num_words"], c["num_entities"], desc, tc)
    rng = nv.RNG(c["seed"])
    m.initialize(rng)
    B, n = c["batch_size"], c["window_size"]
    batch = nv.Batch(B, n).fill(np.full((B, n), c["feature_value"]), np.full(B, c["label"]))
    res = m.compute_cost(batch, rng)
    assert abs(res.get_cost() - 6.17158013374) < 2e-5
    m.compute_gradients(res)
    assert_close(m.get_tensor("grad_transform"), g["grad_transform"], 2e-5, 1e-5)
    assert_close(m.get_tensor("grad_bias"), g["grad_bias"], 2e-5, 1e-5)
    assert_close(m.get_tensor("grad_phrase_reprs"), g["grad_phrase"], 5e-5, 1e-4)


def test_reference_golden_bn_forward():
    """cpp/model_tests.cu:468-521 (batch-norm + tanh forward). The library fixes eps at the
    value the objective uses (1e-4, cpp/objective.cu:114) while that test uses 1e-5, so compare
    with the float64 oracle at 1e-4 instead of the literal vector."""
    desc = nv.ModelDesc(word_repr_size=3, entity_repr_size=5, batch_normalization=True)
    tc = nv.TrainConfig(batch_size=2, window_size=1, num_random_entities=1)
    m = nv.Model(2, 1, desc, tc)
    m.set_tensor(nv.TRANSFORM, np.arange(15)); m.set_tensor(nv.BIAS, np.arange(5) * 1e-3)
    m.set_tensor(nv.WORD_REPRS, [0.01, 0.02, 0.03, 0.001, 0.002, 0.003]); m.set_tensor(nv.ENTITY_REPRS, np.zeros(5))
    om = O.Model(2, 1, 3, 5, batch_normalization=True, bn_epsilon=1e-4)
    om.set("T", np.arange(15.0)); om.set("b", np.arange(5) * 1e-3)
    om.set("W", [0.01, 0.02, 0.03, 0.001, 0.002, 0.003]); om.set("E", np.zeros(5))
    batch = nv.Batch(2, 1).fill([[0], [1]], [0, 0])
    m.compute_cost(batch, entity_ids=[0, 0, 0, 0])
    om.compute_cost([0, 1], [1.0, 1.0], [0, 0, 0, 0], [1.0, 1.0], 1)
    assert_close(m.get_tensor("word_projections"), om.get("Y"), 1e-4)


CASES = {
    # BASELINE.json configs[0]: LSE tanh, batch 4096, |V|=2k |D|=200 d_w=64 d_d=64 z=4
    "C1_lse": dict(V=2000, D=200, dw=64, dd=64, n=10, z=4, B=4096, nonlinearity=nv.TANH, bn=False),
    "lse_bias_neg": dict(V=300, D=1000, dw=128, dd=128, n=10, z=32, B=512, nonlinearity=nv.TANH, bn=False, bias_neg=True),
    # NVSM shape (d_w=300, d_d=256, n=10, z=10, hard_tanh + BN) at a batch the oracle finishes in seconds
    "nvsm_small": dict(V=5000, D=4000, dw=300, dd=256, n=10, z=10, B=2048, nonlinearity=nv.HARD_TANH, bn=True),
    "nvsm_tanh_bn": dict(V=500, D=400, dw=300, dd=256, n=10, z=10, B=1024, nonlinearity=nv.TANH, bn=True),
    # dims that are not multiples of 4 exercise the scalar kernels; ragged batch (not a multiple of anything)
    "odd_dims": dict(V=20, D=15, dw=3, dd=5, n=3, z=1, B=1000, nonlinearity=nv.TANH, bn=True),
    "odd_dims_hard": dict(V=50, D=40, dw=10, dd=6, n=4, z=3, B=777, nonlinearity=nv.HARD_TANH, bn=False),
    "wide": dict(V=100, D=100, dw=512, dd=384, n=2, z=2, B=256, nonlinearity=nv.TANH, bn=False),
    "single_instance": dict(V=10, D=10, dw=8, dd=8, n=1, z=1, B=1, nonlinearity=nv.TANH, bn=False),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_forward_backward_matches_oracle(name):
    c = dict(CASES[name])
    V, D, n, z, B = c["V"], c["D"], c["n"], c["z"], c["B"]
    gm, om, rng = twin_models(**c)
    _, _, cost, ocost = run_forward_backward(gm, om, rng, B, n, V, D, z)
    compare_step_tensors(gm, om, cost, ocost)


TF32_CASES = ["C1_lse", "lse_bias_neg", "nvsm_small", "nvsm_tanh_bn", "wide"]


def rel_fro(actual, expected):
    a = np.asarray(actual, np.float64).ravel(); e = np.asarray(expected, np.float64).ravel()
    return np.linalg.norm(a - e) / max(np.linalg.norm(e), 1e-30)


@pytest.mark.parametrize("name", TF32_CASES)
def test_forward_backward_tf32_tensor_core_gemms(name):
    """NVSM_GEMM_TF32: the three projection GEMMs run on tcgen05 with kind::tf32 (10-bit operand
    mantissas, fp32 accumulation), i.e. ~1e-3 relative error per GEMM output. Smooth quantities are
    compared elementwise at 5e-3; gradients through hard_tanh are compared in relative Frobenius norm
    (2e-2) because a 1e-3 perturbation of a pre-activation that sits on a clip boundary flips that
    element's derivative between 0 and 1. Indices stay bit-exact."""
    c = dict(CASES[name])
    V, D, n, z, B = c["V"], c["D"], c["n"], c["z"], c["B"]
    gm, om, rng = twin_models(**c, gemm_mode=nv.GEMM_TF32)
    _, _, cost, ocost = run_forward_backward(gm, om, rng, B, n, V, D, z)
    assert abs(cost - ocost) <= 2e-3 * abs(ocost)
    assert_close(gm.get_tensor("phrase_reprs"), om.get("P"), 6e-4, what="P")   # stored rounded to tf32 (2^-11)
    assert_close(gm.get_tensor("word_projections"), om.get("Y"), 5e-3, 5e-3, what="Y")
    assert_close(gm.get_tensor("similarity_probs"), om.get("probs"), 5e-3, 2e-3, what="probs")
    for g, o in (("instance_multipliers", "mult"), ("grad_transform", "gT"), ("grad_bias", "gb"),
                 ("grad_phrase_reprs", "gP"), ("grad_entity_repr", "gE")):
        assert rel_fro(gm.get_tensor(g), om.get(o)) <= 2e-2, o


OPTIMISERS = [("sgd", nv.SGD, 0), ("adagrad", nv.ADAGRAD, 0), ("sparse_adam", nv.ADAM, nv.SPARSE),
              ("dense_adam", nv.ADAM, nv.DENSE_UPDATE), ("full_adam", nv.ADAM, nv.DENSE_UPDATE_DENSE_VARIANCE)]
STATE_NAMES = {
    "word_representations-m": "word_m", "word_representations-v": "word_v", "word_representations-acc": "word_acc",
    "entity_representations-m": "entity_m", "entity_representations-v": "entity_v",
    "entity_representations-acc": "entity_acc", "word_entity_mapping-transform-m": "T_m",
    "word_entity_mapping-transform-v": "T_v", "word_entity_mapping-bias-m": "b_m", "word_entity_mapping-bias-v": "b_v",
}


@pytest.mark.parametrize("opt", OPTIMISERS, ids=[o[0] for o in OPTIMISERS])
@pytest.mark.parametrize("bn", [False, True], ids=["nobn", "bn"])
def test_three_training_steps_match_oracle(opt, bn):
    """compute_cost -> compute_gradients -> update, three times, every optimiser of
    cpp/updates*.cu; parameters and optimiser state compared after each step."""
    _, method, mode = opt
    V, D, dw, dd, n, z, B = 300, 200, 32, 24, 5, 3, 512
    lr = 0.01 if method != nv.ADAM else 0.001
    gm, om, rng = twin_models(V, D, dw, dd, n=n, z=z, B=B, nonlinearity=nv.HARD_TANH if bn else nv.TANH, bn=bn,
                              method=method, adam_mode=mode, lam=0.01)
    for step in range(3):
        batch, ids, cost, ocost = run_forward_backward(gm, om, rng, B, n, V, D, z, seed=100 + step)
        assert abs(cost - ocost) <= 5e-4 * abs(ocost)
        lam = gm.L.nvsm_scaled_regularization_lambda(gm.h)
        assert abs(lam - om.scaled_lambda()) <= 1e-7 * abs(lam)
        gm.update(None, lr, lam)
        om.update(lr, om.scaled_lambda())
        for gname, oname in ((nv.WORD_REPRS, "W"), (nv.ENTITY_REPRS, "E"), (nv.TRANSFORM, "T"), (nv.BIAS, "b")):
            assert_close(gm.get_tensor(gname), om.get(oname), 5e-4, 1e-5, what="%s step %d" % (oname, step))
        for gname, oname in STATE_NAMES.items():
            if gm.tensor_size(gname) > 0 and len(om.get(oname)) == gm.tensor_size(gname):
                assert_close(gm.get_tensor(gname), om.get(oname), 1e-3, 1e-5, what="%s step %d" % (oname, step))


@pytest.mark.parametrize("opt", [("sgd", nv.SGD, 0), ("adagrad", nv.ADAGRAD, 0), ("full_adam", nv.ADAM, nv.DENSE_UPDATE_DENSE_VARIANCE)],
                         ids=["sgd", "adagrad", "full_adam"])
def test_zipf_skewed_ids_take_the_heavy_row_path(opt):
    """Zipf(1) word ids and Zipf(1) negatives: the most frequent word / document of a batch is referenced ~1000 times,
    so the pull-style updates split those rows into 64-reference segments (pull_heavy_kernel: per-segment partial sums,
    last arrival applies the row update). Three steps against the oracle, tables and optimiser state."""
    _, method, mode = opt
    V, D, dw, dd, n, z, B = 9000, 8500, 24, 20, 5, 4, 2048       # >= 8192 rows: SGD / Adagrad pull as well
    lr = 0.01 if method != nv.ADAM else 0.001
    gm, om, rng = twin_models(V, D, dw, dd, n=n, z=z, B=B, nonlinearity=nv.HARD_TANH, bn=True, method=method, adam_mode=mode, lam=0.01)
    gm.set_negative_distribution(nv.zipf_cdf(D, 1.0))
    wcdf = nv.zipf_cdf(V, 1.0)
    nrng = np.random.default_rng(4)
    for step in range(3):
        f, fw, labels, w = make_batch(nrng, B, n, V, D, z)
        f = np.minimum(np.searchsorted(wcdf, nrng.random((B, n)), side="right"), V - 1).astype(np.int64)
        ids = gm.generate_labels(labels, rng)
        assert np.bincount(f.ravel()).max() > 500 and np.bincount(ids).max() > 300     # the heavy path is really taken
        batch = nv.Batch(B, n).fill(f, labels, fw, w)
        res = gm.compute_cost(batch, entity_ids=ids)
        cost, ocost = res.get_cost(), om.compute_cost(f, fw, ids, w, n)
        assert abs(cost - ocost) <= 5e-4 * abs(ocost)
        gm.compute_gradients(res); om.compute_gradients()
        gm.update(None, lr, res.scaled_regularization_lambda()); om.update(lr, om.scaled_lambda())
        for gname, oname in ((nv.WORD_REPRS, "W"), (nv.ENTITY_REPRS, "E"), (nv.TRANSFORM, "T"), (nv.BIAS, "b")):
            assert_close(gm.get_tensor(gname), om.get(oname), 5e-4, 1e-5, what="%s step %d" % (oname, step))
        for gname, oname in STATE_NAMES.items():
            if gm.tensor_size(gname) > 0 and len(om.get(oname)) == gm.tensor_size(gname):
                assert_close(gm.get_tensor(gname), om.get(oname), 1e-3, 1e-5, what="%s step %d" % (oname, step))


def test_unweighted_rebalanced_lse_and_duplicates():
    """z > 1 without bias_negative_samples re-weights positives/negatives
    (cpp/objective.cu:268-290); tiny D forces id collisions (negatives == positive, duplicate rows
    inside one n-gram) so the scatter must accumulate."""
    c = dict(V=7, D=3, dw=16, dd=16, n=6, z=5, B=2048, nonlinearity=nv.TANH, bn=False, bias_neg=False, method=nv.SGD)
    gm, om, rng = twin_models(**c)
    batch, ids, cost, ocost = run_forward_backward(gm, om, rng, c["B"], c["n"], c["V"], c["D"], c["z"], weighted=False)
    compare_step_tensors(gm, om, cost, ocost)
    lam = gm.L.nvsm_scaled_regularization_lambda(gm.h)
    gm.update(None, 0.05, lam); om.update(0.05, om.scaled_lambda())
    assert_close(gm.get_tensor(nv.ENTITY_REPRS), om.get("E"), 5e-4, 1e-5)
    assert_close(gm.get_tensor(nv.WORD_REPRS), om.get("W"), 5e-4, 1e-5)


def test_clip_sigmoid_saturation_band():
    """Large scores saturate the sigmoid: the forward clamp (1e-7) and the backward zero band
    (1e-6) must agree with the oracle (cpp/objective.cu:242-246,354-371)."""
    c = dict(V=10, D=10, dw=8, dd=8, n=2, z=2, B=64, nonlinearity=nv.HARD_TANH, bn=False)
    gm, om, rng = twin_models(**c)
    big = np.random.default_rng(0).uniform(-30, 30, size=10 * 8).astype(np.float32)
    gm.set_tensor(nv.ENTITY_REPRS, big); om.set("E", big)
    T = np.random.default_rng(1).uniform(-5, 5, size=64).astype(np.float32)
    gm.set_tensor(nv.TRANSFORM, T); om.set("T", T)
    _, _, cost, ocost = run_forward_backward(gm, om, rng, 64, 2, 10, 10, 2)
    probs = gm.get_tensor("similarity_probs")
    assert probs.min() >= np.float32(1e-7) and probs.max() <= 1.0
    assert (gm.get_tensor("instance_multipliers") == 0).sum() > 0
    compare_step_tensors(gm, om, cost, ocost)


def test_infer_matches_oracle():
    gm, om, rng = twin_models(200, 50, 64, 32, n=4, z=1, B=16, nonlinearity=nv.HARD_TANH, bn=True)
    words = np.random.default_rng(5).integers(0, 200, size=(33, 4))
    assert_close(gm.infer(words, 4), om.infer(words, 4), RTOL)


def test_staged_batch_equals_host_batch():
    c = dict(V=500, D=300, dw=64, dd=64, n=5, z=4, B=1024, nonlinearity=nv.TANH, bn=True)
    gm, om, rng = twin_models(**c, num_batch_slots=2)
    f, fw, labels, w = make_batch(np.random.default_rng(9), 1024, 5, 500, 300, 4)
    batch = nv.Batch(1024, 5).fill(f, labels, fw, w)
    ids = gm.generate_labels(labels, rng)
    a = gm.compute_cost(batch, entity_ids=ids).get_cost()
    gm.stage_batch(1, batch, ids)
    b = gm.compute_cost_staged(1).get_cost()
    assert abs(a - b) <= 1e-6 * abs(a)


def test_null_weights_mean_uniform_weights_bit_exact():
    """feature_weights / weights == NULL over the C ABI (Batch.fill without weights) is the reference's uniform weighting:
    the device copy is filled with ones instead of being transferred. Same results as passing arrays of 1.0 through
    every upload path, also when the batch slot held other weights before (two model instances: float atomics and the
    bucket order make later steps agree to round-off, not bitwise)."""
    c = dict(V=600, D=400, dw=300, dd=256, n=10, z=10, B=512, nonlinearity=nv.TANH, bn=True,   # (tanh: no clip-boundary flips)
             method=nv.ADAM, adam_mode=nv.DENSE_UPDATE_DENSE_VARIANCE)
    a, _, rng_a = twin_models(**c, num_batch_slots=2)
    b, _, rng_b = twin_models(**c, num_batch_slots=2)
    nrng = np.random.default_rng(21)
    ones_fw, ones_w = np.ones((512, 10), np.float32), np.ones(512, np.float32)
    for step in range(4):
        f, fw, labels, w = make_batch(nrng, 512, 10, 600, 400, 10)
        ids = a.generate_labels(labels, rng_a)
        assert (b.generate_labels(labels, rng_b) == ids).all()
        if step == 1:   # a weighted batch in between: the slots' "holds ones" state must be invalidated
            explicit, implicit = nv.Batch(512, 10).fill(f, labels, fw, w), nv.Batch(512, 10).fill(f, labels, fw, w)
        else:
            explicit, implicit = nv.Batch(512, 10).fill(f, labels, ones_fw, ones_w), nv.Batch(512, 10).fill(f, labels)
            assert implicit.uniform_feature_weights_ and implicit.uniform_weights_ and not explicit.uniform_weights_
        ra, rb = a.compute_cost(explicit, entity_ids=ids), b.compute_cost(implicit, entity_ids=ids)
        ca, cb = ra.get_cost(), rb.get_cost()
        # (round-off, not bitwise, from the first step on: the per-block loss is combined with float atomics in shared
        # memory, whose order follows the warps' finishing order)
        assert abs(ca - cb) <= 2e-6 * abs(ca), (step, ca, cb)
        a.backprop(ra, 0.01); b.backprop(rb, 0.01)
        a.train_step(explicit, ids, 0.01); b.train_step(implicit, ids, 0.01)        # copy-stream upload path
        a.stage_batch(1, explicit, ids); b.stage_batch(1, implicit, ids)            # staged path
        a.train_step_staged(1, 0.01); b.train_step_staged(1, 0.01)
        assert abs(a.last_cost() - b.last_cost()) <= 2e-6 * abs(a.last_cost()), step
    for name in (nv.WORD_REPRS, nv.ENTITY_REPRS, nv.TRANSFORM, nv.BIAS):
        assert_close(a.get_tensor(name), b.get_tensor(name), 2e-4, 2e-4, name)


def test_errors_are_reported_not_swallowed():
    with pytest.raises(nv.NvsmError):   # a mixture objective with a zero weight (CHECK_NE in cpp/objective.cu:709-710)
        nv.Model(10, 10, nv.ModelDesc(), nv.TrainConfig(text_entity_weight=1.0, entity_entity_weight=0.0),
                 objective=nv.TEXT_ENTITY_ENTITY_ENTITY)
    with pytest.raises(nv.NvsmError):   # a pair forward on a plain TextEntity handle
        nv.Model(10, 10, nv.ModelDesc(), nv.TrainConfig(batch_size=8, window_size=2)).similarity_compute_cost(
            nv.SimilarityBatch(4).fill(np.zeros((4, 2))))
    m = nv.Model(10, 10, nv.ModelDesc(), nv.TrainConfig(batch_size=8, window_size=2))
    with pytest.raises(nv.NvsmError):
        m.compute_gradients()          # no forward result
    with pytest.raises(nv.NvsmError):
        m.get_tensor("word_representations-representations"[:-1] + "x") if m.tensor_size("nope") >= 0 else m.update(None, 0.1, 0.0)
    batch = nv.Batch(16, 2).fill(np.zeros((16, 2)), np.zeros(16))
    with pytest.raises(nv.NvsmError):
        m.compute_cost(batch, nv.RNG(1))  # larger than max_batch_size


def test_out_of_range_ids_are_reported_and_cannot_corrupt_the_tables():
    """The reference only DCHECKs its ids; here a batch with ids outside the tables is clamped on the device behind
    its upload and reported by the next synchronising call, on every upload path. The handle stays usable."""
    c = dict(V=300, D=200, dw=32, dd=32, n=4, z=3, B=256, nonlinearity=nv.TANH, bn=False, num_batch_slots=2)
    gm, om, rng = twin_models(**c)
    f, fw, labels, w = make_batch(np.random.default_rng(2), c["B"], c["n"], c["V"], c["D"], c["z"])
    W0, E0 = gm.get_tensor(nv.WORD_REPRS).copy(), gm.get_tensor(nv.ENTITY_REPRS).copy()
    ids = gm.generate_labels(labels, rng)

    bad_f = f.copy(); bad_f[17, 2] = c["V"]                       # one past the word table
    res = gm.compute_cost(nv.Batch(c["B"], c["n"]).fill(bad_f, labels, fw, w), entity_ids=ids)
    with pytest.raises(nv.NvsmError, match="word ids outside"):
        res.get_cost()
    bad_ids = ids.copy(); bad_ids[5] = -1                         # negative entity id
    with pytest.raises(nv.NvsmError, match="entity ids outside"):
        gm.stage_batch(1, nv.Batch(c["B"], c["n"]).fill(f, labels, fw, w), bad_ids)
    bad_labels = labels.copy(); bad_labels[3] = 10 ** 12          # device-sampler path: the positive label
    gm.sampler_seed(nv.RNG(5))
    gm.step_sampled(nv.Batch(c["B"], c["n"]).fill(f, bad_labels, fw, w), 0.01)
    with pytest.raises(nv.NvsmError, match="entity ids outside"):
        gm.last_cost()
    gm.synchronize()                                              # reported once, then cleared
    assert np.isfinite(gm.get_tensor(nv.WORD_REPRS)).all() and np.isfinite(gm.get_tensor(nv.ENTITY_REPRS)).all()

    # a clean batch afterwards behaves like a fresh model would (parameters restored)
    gm.set_tensor(nv.WORD_REPRS, W0); gm.set_tensor(nv.ENTITY_REPRS, E0)
    gm.set_tensor(nv.TRANSFORM, om.get("T")); gm.set_tensor(nv.BIAS, om.get("b"))
    cost = gm.compute_cost(nv.Batch(c["B"], c["n"]).fill(f, labels, fw, w), entity_ids=ids).get_cost()
    ocost = om.compute_cost(f, fw, ids, w, c["n"])
    assert abs(cost - ocost) <= RTOL * abs(ocost)


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[1] at full size: size-independent properties.
# ---------------------------------------------------------------------------------------------
def _c2_model(method=nv.SGD, mode=0, lam=0.0):
    desc = nv.ModelDesc(word_repr_size=300, entity_repr_size=256, batch_normalization=True,
                        nonlinearity=nv.HARD_TANH, clip_sigmoid=True)
    tc = nv.TrainConfig(batch_size=51200, window_size=10, num_random_entities=10, regularization_lambda=lam,
                        update_method=method, adam_mode=mode)
    m = nv.Model(50000, 50000, desc, tc)
    m.initialize(nv.RNG(1))
    return m


def test_full_size_c2_properties():
    B, n, z, V, D = 51200, 10, 10, 50000, 50000
    m = _c2_model()
    nrng = np.random.default_rng(11)
    f, fw, labels, w = make_batch(nrng, B, n, V, D, z, weighted=False)
    rng = nv.RNG(12345)
    ids = m.generate_labels(labels, rng)
    batch = nv.Batch(B, n).fill(f, labels, fw, w)
    res = m.compute_cost(batch, entity_ids=ids)
    cost = res.get_cost()
    assert np.isfinite(cost) and 0.0 < cost < 50.0
    m.compute_gradients(res)
    gT, gb, gP = m.get_tensor("grad_transform"), m.get_tensor("grad_bias"), m.get_tensor("grad_phrase_reprs")
    mult = m.get_tensor("instance_multipliers")
    # (1) batch-norm backward: the batch sum of d cost / d pre-activation vanishes per feature, so
    #     grad_phrase summed over the batch is ~0 (sum_i dX_i = 0 => sum_i T^T dX_i = 0).
    gp_sum = gP.reshape(B, 300).sum(0)
    assert np.abs(gp_sum).max() <= 1e-3 * np.abs(gP).sum() / 300 + 1e-7
    # (2) permutation invariance: shuffling the n-grams changes only summation order.
    perm = nrng.permutation(B)
    b2 = nv.Batch(B, n).fill(f[perm], labels[perm], fw[perm], w[perm])
    ids2 = ids.reshape(B, z + 1)[perm].ravel()
    res2 = m.compute_cost(b2, entity_ids=ids2)
    assert abs(res2.get_cost() - cost) <= 1e-5 * abs(cost)
    m.compute_gradients(res2)
    assert_close(m.get_tensor("grad_transform"), gT, 1e-3, 1e-4)
    assert_close(m.get_tensor("grad_bias"), gb, 1e-3, 1e-3)
    assert_close(m.get_tensor("instance_multipliers").reshape(B, z + 1)[np.argsort(perm)], mult.reshape(B, z + 1), 1e-3, 1e-5)
    # (3) checksum of the sparse SGD update (lambda = 0): the column sums of the table delta equal
    #     the column sums of lr * sum_c grad_entity[c] / lr * sum_{i,w} fw * grad_phrase[i].
    E0 = m.get_tensor(nv.ENTITY_REPRS).astype(np.float64).reshape(D, 256)
    W0 = m.get_tensor(nv.WORD_REPRS).astype(np.float64).reshape(V, 300)
    Y = m.get_tensor("word_projections").astype(np.float64).reshape(B, 256)
    mult2 = m.get_tensor("instance_multipliers").astype(np.float64).reshape(B, z + 1)
    gP2 = m.get_tensor("grad_phrase_reprs").astype(np.float64).reshape(B, 300)
    sign = np.where(np.arange(z + 1) == 0, 1.0, -1.0)
    lr = 0.5
    m.update(None, lr, 0.0)
    dE = m.get_tensor(nv.ENTITY_REPRS).astype(np.float64).reshape(D, 256) - E0
    dW = m.get_tensor(nv.WORD_REPRS).astype(np.float64).reshape(V, 300) - W0
    expect_E = lr * ((mult2 * sign).sum(1)[:, None] * Y).sum(0)
    expect_W = lr * n * gP2.sum(0)
    # tolerance relative to the total mass that was scattered into each column (the signed column
    # sums themselves nearly cancel under batch-norm)
    mass_E = lr * (np.abs(mult2).sum(1)[:, None] * np.abs(Y)).sum(0)
    mass_W = lr * n * np.abs(gP2).sum(0)
    assert (np.abs(dE.sum(0) - expect_E) <= 1e-4 * mass_E + 1e-9).all()
    assert (np.abs(dW.sum(0) - expect_W) <= 1e-4 * mass_W + 1e-9).all()
    # rows never referenced stay bit-identical
    untouched = np.setdiff1d(np.arange(D), np.unique(ids2))
    assert (dE[untouched] == 0).all()


def test_full_size_c2_full_adam_runs_and_learns():
    """A few full_adam steps at the headline configuration: finite, and the loss goes down."""
    B, n, z, V, D = 51200, 10, 10, 50000, 50000
    m = _c2_model(nv.ADAM, nv.DENSE_UPDATE_DENSE_VARIANCE, lam=0.01)
    nrng = np.random.default_rng(2)
    f, fw, labels, w = make_batch(nrng, B, n, V, D, z, weighted=False)
    labels = f[:, 0] % D  # learnable signal: the document is a function of the first word
    batch = nv.Batch(B, n).fill(f, labels, fw, w)
    rng = nv.RNG(99)
    costs = []
    for step in range(8):
        ids = m.generate_labels(labels, rng)
        m.train_step(batch, ids, 0.001)
        costs.append(m.last_cost())
    assert np.isfinite(costs).all()
    assert costs[-1] < costs[0]


# ---------------------------------------------------------------------------------------------
# Device-side sampler: bit-exact with the reference's host loop (cpp/labels.cu:3-22)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D,z,B,seed", [(3, 10, 32, 10), (50000, 10, 51200, 1), (1000000, 32, 4096, 77), (200, 4, 4096, 5),
                                        (7, 0, 100, 4), (1, 5, 64, 2)])
def test_device_sampler_bit_exact(D, z, B, seed):
    """ids and engine state after three consecutive batches (engine state kept on the device) equal the
    oracle's serial std::uniform_int_distribution<long> draws."""
    n = 2
    desc = nv.ModelDesc(word_repr_size=4, entity_repr_size=4, clip_sigmoid=True)
    tc = nv.TrainConfig(batch_size=B, window_size=n, num_random_entities=z)
    m = nv.Model(10, D, desc, tc)
    rng = nv.RNG(seed)
    m.sampler_seed(rng)
    state = rng.state
    nrng = np.random.default_rng(seed)
    for step in range(3):
        labels = nrng.integers(0, D, size=B, dtype=np.int64)
        batch = nv.Batch(B, n).fill(np.zeros((B, n)), labels)
        m.step_sampled(batch, 0.0, train=False)
        got = m.entity_ids(B)
        exp, state = O.generate_labels(labels, z, D, state)
        assert (got == exp).all(), "step %d" % step
        assert m.sampler_state() == state


@pytest.mark.parametrize("D,z,B,seed", [(1 << 30, 3, 20000, 9), (1500000000, 2, 40000, 3), (2147483645, 1, 30000, 11),
                                        (1073741825, 4, 10000, 6), (50000, 10, 51200, 2)])
def test_device_sampler_large_range(D, z, B, seed):
    """Stand-alone device sampling for ranges where libstdc++'s rejection loop fires often
    (D = 1.5e9 rejects ~30%, D = 2^30 + 1 rejects ~50% of the candidates)."""
    m = nv.Model(4, 4, nv.ModelDesc(word_repr_size=4, entity_repr_size=4), nv.TrainConfig(batch_size=8, window_size=1))
    labels = np.random.default_rng(seed).integers(0, D, size=B, dtype=np.int64)
    rng = nv.RNG(seed)
    got = m.generate_labels_device(labels, rng, z=z, num_objects=D)
    exp, state = O.generate_labels(labels, z, D, seed)
    assert (got == exp).all()
    assert rng.state == state


def test_device_sampled_step_equals_host_sampled_step():
    """nvsm_step_sampled == nvsm_generate_labels + nvsm_train_step (same ids => same costs, same tables)."""
    c = dict(V=400, D=300, dw=32, dd=32, n=4, z=6, B=1024, nonlinearity=nv.HARD_TANH, bn=True,
             method=nv.ADAM, adam_mode=nv.DENSE_UPDATE_DENSE_VARIANCE)
    a, _, rng_a = twin_models(**c)
    b, _, rng_b = twin_models(**c)
    assert rng_a.state == rng_b.state
    b.sampler_seed(rng_b)
    nrng = np.random.default_rng(1)
    for step in range(4):
        f, fw, labels, w = make_batch(nrng, c["B"], c["n"], c["V"], c["D"], c["z"])
        batch = nv.Batch(c["B"], c["n"]).fill(f, labels, fw, w)
        ids = a.generate_labels(labels, rng_a)
        a.train_step(batch, ids, 0.001)
        b.step_sampled(batch, 0.001)
        assert (b.entity_ids(c["B"]) == ids).all()
        assert abs(a.last_cost() - b.last_cost()) <= 1e-6 * abs(a.last_cost())
    assert b.sampler_state() == rng_a.state
    assert_close(b.get_tensor(nv.ENTITY_REPRS), a.get_tensor(nv.ENTITY_REPRS), 1e-5, 1e-6)


def test_timeline_mode_reports_overlapped_phases_and_does_not_change_results():
    """nvsm_set_profiling(m, 2): the stream overlaps of the fused step stay on, every phase reports (start, end) against one
    origin; the step's results are the ones of an unprofiled model (same kernels, same order per table)."""
    V, D, dw, dd, n, z, B = 9000, 9000, 300, 256, 10, 10, 4096
    kw = dict(n=n, z=z, B=B, nonlinearity=nv.TANH, bn=True, method=nv.ADAM, adam_mode=nv.DENSE_UPDATE_DENSE_VARIANCE,
              gemm_mode=nv.GEMM_3XTF32)   # (tanh: no clip boundary for round-off to flip)
    a, _, rng_a = twin_models(V, D, dw, dd, **kw)
    b, _, rng_b = twin_models(V, D, dw, dd, **kw)
    f, fw, labels, w = make_batch(np.random.default_rng(5), B, n, V, D, z)
    ids = a.generate_labels(labels, rng_a)
    batch = nv.Batch(B, n).fill(f, labels, fw, w)
    b.set_profiling(2)
    for m in (a, b):
        m.train_step(batch, ids, 0.01)
        m.train_step(batch, ids, 0.01)
    tl = b.timeline()
    b.set_profiling(0)
    names = [t[0] for t in tl]
    for ph in ("gather_mean", "gemm_fwd", "score_loss_bwd", "bn_backward", "gemm_grad_transform", "gemm_grad_phrase",
               "update_entities", "update_words", "update_transform", "bucket_build"):
        assert names.count(ph) == 2, (ph, names)
    assert all(0.0 <= s <= e for _, s, e in tl)
    ent = [t for t in tl if t[0] == "update_entities"][0]
    gp = [t for t in tl if t[0] == "gemm_grad_phrase"][0]
    assert ent[1] < gp[2], "the entity update runs on the auxiliary stream under the backward GEMMs"
    # (not bit-for-bit: the order of a row's references inside its bucket follows the arrival order of integer atomics)
    assert abs(a.last_cost() - b.last_cost()) <= 1e-6 * abs(a.last_cost())
    for name in (nv.WORD_REPRS, nv.ENTITY_REPRS, nv.TRANSFORM, nv.BIAS):
        assert_close(b.get_tensor(name), a.get_tensor(name), 1e-5, 1e-6, name)
    a.close(); b.close()


def test_programmatic_dependent_launch_mode_keeps_parity():
    """NVSM_PDL=63 (every launch site as a programmatic dependent of the kernel in front of it) in a fresh process: the
    tensor-core forward / backward and three-step training cases still match the oracle."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, NVSM_PDL="63")
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q", "-x",
                          "-k", "tf32_tensor_core_gemms or three_training_steps or full_size_c2_full_adam"],
                         env=env, capture_output=True, text=True, timeout=900, cwd=root)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
