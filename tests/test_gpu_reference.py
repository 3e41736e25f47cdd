"""-m gpu: parity against the REFERENCE ITSELF, run here: oracle/_ref/libcunvsm_ref_{f32,f64}.so is the
unmodified cpp/{model,objective,params,storage,updates*,intermediate_results,labels,cudnn_utils,
cuda_utils,data}.cu of /root/reference compiled against oracle/ref_shim (a reconstruction of the
un-vendored device_matrix, cuBLAS + cuDNN from the image).

Three links of the chain are checked:
  1. shim validity: the reference's own known-answer test Transform_backward
     (cpp/model_tests.cu:341-466) is replayed through the float64 reference build and must reproduce
     the literals it asserts (DoubleEq => 4 ulp; we allow 1e-12 relative).
  2. oracle == reference: the float64 CPU oracle and the float64 reference on seeded inputs, every
     flag / optimiser combination, three optimiser steps: rtol 1e-9 (summation order only).
  3. sm_100a path == reference: libnvsm_b200.so against the float32 -use_fast_math reference (the
     release library): ids and RNG state bit-exact, Glorot init bit-exact, fp32 tolerance 2e-4
     relative (+1e-6 x max floor) for step tensors with the fp32-SIMT GEMM, 5e-4 after updates.
The tests skip when oracle/_ref was not built (it can only be built where /root/reference exists).
"""
import json
import os

import numpy as np
import pytest

import cunvsm_b200 as nv
from oracle import binding as O
from oracle import ref_binding as R
from tests.util import assert_close, make_batch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (R.available(np.float32) and R.available(np.float64)),
                                 reason="oracle/_ref not built (needs /root/reference)")]
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_goldens.json")))

CASES = [
    # V, D, dw, dd, n, z, B, nonlinearity, bn, clip, bias_neg
    dict(V=50, D=40, dw=8, dd=6, n=3, z=2, B=64, nonlinearity=nv.TANH, bn=False, clip=True, bias_neg=False),
    dict(V=300, D=200, dw=64, dd=64, n=10, z=4, B=256, nonlinearity=nv.TANH, bn=False, clip=True, bias_neg=True),
    dict(V=500, D=400, dw=300, dd=256, n=10, z=10, B=512, nonlinearity=nv.HARD_TANH, bn=True, clip=True,
         bias_neg=False),
    dict(V=120, D=90, dw=20, dd=12, n=4, z=1, B=128, nonlinearity=nv.HARD_TANH, bn=True, clip=False, bias_neg=False),
]
METHODS = [(nv.SGD, 0), (nv.ADAGRAD, 0), (nv.ADAM, nv.SPARSE), (nv.ADAM, nv.DENSE_UPDATE),
           (nv.ADAM, nv.DENSE_UPDATE_DENSE_VARIANCE)]
PARAMS = (("word_representations", "W", nv.WORD_REPRS), ("entity_representations", "E", nv.ENTITY_REPRS),
          ("transform", "T", nv.TRANSFORM), ("bias", "b", nv.BIAS))


def ref_model(c, method, dtype, lam=0.01, seed=7):
    return R.Model(c["V"], c["D"], c["dw"], c["dd"], batch_size=c["B"], window_size=c["n"],
                   num_random_entities=c["z"], nonlinearity=c["nonlinearity"], batch_normalization=c["bn"],
                   clip_sigmoid=c["clip"], bias_negative_samples=c["bias_neg"], update_method=method[0],
                   l2_normalize_phrase_reprs=c.get("l2_phrase", False), l2_normalize_entity_reprs=c.get("l2_entity", False),
                   adam_mode=method[1], regularization_lambda=lam, seed=seed, dtype=dtype)


def oracle_model(c, method, dtype, lam=0.01):
    return O.Model(c["V"], c["D"], c["dw"], c["dd"], nonlinearity=c["nonlinearity"], batch_normalization=c["bn"],
                   clip_sigmoid=c["clip"], bias_negative_samples=c["bias_neg"], update_method=method[0],
                   adam_mode=method[1], num_random_entities=c["z"], regularization_lambda=float(np.float32(lam)),
                   dtype=dtype)


def cuda_model(c, method, lam=0.01, gemm_mode=nv.GEMM_FP32):
    desc = nv.ModelDesc(word_repr_size=c["dw"], entity_repr_size=c["dd"], batch_normalization=c["bn"],
                        nonlinearity=c["nonlinearity"], clip_sigmoid=c["clip"], bias_negative_samples=c["bias_neg"],
                        l2_normalize_phrase_reprs=c.get("l2_phrase", False),
                        l2_normalize_entity_reprs=c.get("l2_entity", False))
    tc = nv.TrainConfig(batch_size=c["B"], window_size=c["n"], num_random_entities=c["z"],
                        regularization_lambda=lam, update_method=method[0], adam_mode=method[1])
    return nv.Model(c["V"], c["D"], desc, tc, gemm_mode=gemm_mode)


def test_shim_reproduces_reference_known_answer_test():
    """cpp/model_tests.cu:341-466 replayed through the float64 reference build."""
    g = GOLD["transform_backward"]; c = g["config"]
    B, n, z = c["batch_size"], c["window_size"], c["num_random_entities"]
    rm = R.Model(c["num_words"], c["num_entities"], c["word_repr_size"], c["entity_repr_size"], batch_size=B,
                 window_size=n, num_random_entities=z, nonlinearity=R.TANH, bias_negative_samples=True,
                 seed=c["seed"], dtype=np.float64)
    batch = rm.new_batch().fill(np.full((B, n), c["feature_value"]), np.full(B, c["label"]), np.ones((B, n)), np.ones(B))
    rm.forward(batch)
    assert abs(rm.get_cost() - 6.17158013374) < 1e-9  # value printed by the survey's probe
    rm.compute_gradients()
    np.testing.assert_allclose(rm.get("grad_transform"), np.array(g["grad_transform"]), rtol=1e-12, atol=0)
    np.testing.assert_allclose(rm.get("grad_bias"), np.array(g["grad_bias"]), rtol=1e-12, atol=0)
    np.testing.assert_allclose(rm.get("grad_phrase_reprs"), np.array(g["grad_phrase"]), rtol=1e-12, atol=0)


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("method", METHODS, ids=lambda m: "m%d_%d" % m)
def test_oracle_matches_reference_float64(case, method):
    c = CASES[case]
    rm = ref_model(c, method, np.float64)
    om = oracle_model(c, method, np.float64)
    # Glorot init: the oracle consumes the engine exactly like Model::initialize (W, E, T; b = 0).
    state = om.initialize(7)
    assert state == rm.rng_state
    for rname, oname, _ in PARAMS:
        np.testing.assert_array_equal(rm.get(rname), om.get(oname), err_msg="init " + rname)
    nrng = np.random.default_rng(11 + case)
    lr = 0.05
    for step in range(3):
        f, fw, labels, w = make_batch(nrng, c["B"], c["n"], c["V"], c["D"], c["z"])
        ids, state = O.generate_labels(labels, c["z"], c["D"], state)
        rm.forward(rm.new_batch().fill(f, labels, fw, w))
        assert (rm.entity_ids() == ids).all() and rm.rng_state == state
        ocost = om.compute_cost(f, fw.astype(np.float64), ids, w.astype(np.float64), c["n"])
        assert abs(rm.get_cost() - ocost) <= 1e-10 * abs(ocost)
        rm.compute_gradients(); om.compute_gradients()
        for rname, oname in (("phrase_reprs", "P"), ("word_projections", "Y"), ("similarity_probs", "probs"),
                             ("grad_transform", "gT"), ("grad_bias", "gb"), ("grad_phrase_reprs", "gP"),
                             ("grad_entity_repr", "gE")):
            assert_close(rm.get(rname), om.get(oname), 1e-9, 1e-12, "%s step %d" % (rname, step))
        assert abs(rm.scaled_lambda() - om.scaled_lambda()) <= 1e-15
        rm.update(lr, rm.scaled_lambda()); om.update(lr, om.scaled_lambda())
        for rname, oname, _ in PARAMS:
            assert_close(rm.get(rname), om.get(oname), 1e-9, 1e-12, "%s after step %d" % (rname, step))


def compare_cuda_with_reference(c, method, seed_batches, gemm_mode=nv.GEMM_FP32, rtol=2e-4):
    rm = ref_model(c, method, np.float32)
    gm = cuda_model(c, method, gemm_mode=gemm_mode)
    rng = nv.RNG(7)
    gm.initialize(rng)
    assert rng.state == rm.rng_state
    for rname, _, gname in PARAMS:
        np.testing.assert_array_equal(rm.get(rname), gm.get_tensor(gname), err_msg="init " + rname)
    nrng = np.random.default_rng(seed_batches)
    lr = 0.05
    for step in range(3):
        f, fw, labels, w = make_batch(nrng, c["B"], c["n"], c["V"], c["D"], c["z"])
        rm.forward(rm.new_batch().fill(f, labels, fw, w))
        res = gm.compute_cost(nv.Batch(c["B"], c["n"]).fill(f, labels, fw, w), rng)
        assert (gm.entity_ids(c["B"]) == rm.entity_ids()).all(), "sampled ids bit-exact"
        assert rng.state == rm.rng_state
        rcost = rm.get_cost()
        assert abs(res.get_cost() - rcost) <= rtol * abs(rcost) + 1e-7
        rm.compute_gradients(); gm.compute_gradients(res)
        for name, floor in (("phrase_reprs", 1e-6), ("word_projections", 2e-6), ("similarity_probs", 1e-6),
                            ("grad_transform", 1e-5), ("grad_bias", 1e-5), ("grad_phrase_reprs", 1e-5),
                            ("grad_entity_repr", 1e-5)):
            assert_close(gm.get_tensor(name), rm.get(name), rtol, floor, "%s step %d" % (name, step))
        rm.update(lr, rm.scaled_lambda()); gm.update(None, lr, res.scaled_regularization_lambda())
        for rname, _, gname in PARAMS:
            ref_value = rm.get(rname)
            # Adam steps are ~lr whatever the gradient's size (m / sqrt(v)): rounding noise on near-zero gradients is
            # amplified to a fraction of lr, hence the absolute floor of 2e-3 x lr for the Adam modes.
            floor = 1e-5 if method[0] != nv.ADAM else 2e-3 * lr / max(np.abs(ref_value).max(), 1e-30)
            assert_close(gm.get_tensor(gname), ref_value, 5e-4, max(floor, 1e-5), "%s after step %d" % (rname, step))
            # re-align the parameters (not the optimiser state) so that every step is checked at the per-step
            # tolerance; the free-running trajectory is the subject of the next test and test_gpu_loss_curve.py
            gm.set_tensor(gname, ref_value)


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("method", METHODS, ids=lambda m: "m%d_%d" % m)
def test_cuda_path_matches_reference_float32(case, method):
    compare_cuda_with_reference(CASES[case], method, 23 + case)


@pytest.mark.parametrize("l2", [(True, False), (False, True), (True, True)], ids=["phrase", "entity", "both"])
@pytest.mark.parametrize("method", METHODS, ids=lambda m: "m%d_%d" % m)
@pytest.mark.parametrize("case", [1, 2])
def test_l2_normalizer_matches_reference_float32(case, method, l2):
    """--l2_phrase_normalization / --l2_entity_normalization (Normalizer fwd/bwd, cpp/cuda_utils.cu:3-141) against the
    reference for every optimiser: LSE shape without batch-norm and the NVSM shape with batch-norm + hard_tanh."""
    c = dict(CASES[case], l2_phrase=l2[0], l2_entity=l2[1])
    compare_cuda_with_reference(c, method, 31 + case)


def test_cuda_tensor_core_path_matches_reference_nvsm_shape():
    """The default bench arithmetic (3xTF32 tcgen05 GEMMs, ring score kernel, pull Adam) against the float32
    reference on the NVSM shape (d_w=300, d_d=256, n=10, z=10, BN + hard_tanh, full Adam), B=2048, 5 steps."""
    c = dict(V=3000, D=2500, dw=300, dd=256, n=10, z=10, B=2048, nonlinearity=nv.HARD_TANH, bn=True, clip=True,
             bias_neg=False)
    method = (nv.ADAM, nv.DENSE_UPDATE_DENSE_VARIANCE)
    rm = ref_model(c, method, np.float32)
    gm = cuda_model(c, method, gemm_mode=nv.GEMM_3XTF32)
    rng = nv.RNG(7)
    gm.initialize(rng)
    theta0 = {gname: gm.get_tensor(gname).astype(np.float64) for _, _, gname in PARAMS}
    nrng = np.random.default_rng(5)
    for step in range(5):
        f, fw, labels, w = make_batch(nrng, c["B"], c["n"], c["V"], c["D"], c["z"], weighted=False)
        rcost = rm.step(rm.new_batch().fill(f, labels, fw, w), 1e-3)
        res = gm.compute_cost(nv.Batch(c["B"], c["n"]).fill(f, labels, fw, w), rng)
        gm.backprop(res, 1e-3)
        assert rng.state == rm.rng_state
        assert abs(res.get_cost() - rcost) <= 2e-4 * abs(rcost), (step, res.get_cost(), rcost)
    # Five free-running full-Adam steps move every element by up to 5 x lr; compare the accumulated update: relative
    # Frobenius error 1e-2 (Adam amplifies rounding noise on near-zero gradients to a fraction of lr), elementwise
    # within 5 % of the total possible movement.
    for rname, _, gname in PARAMS:
        delta_ref = rm.get(rname).astype(np.float64) - theta0[gname]
        delta = gm.get_tensor(gname).astype(np.float64) - theta0[gname]
        err = np.linalg.norm(delta - delta_ref) / max(np.linalg.norm(delta_ref), 1e-30)
        assert err <= 1e-2, (rname, err)
        assert np.abs(delta - delta_ref).max() <= 0.05 * 5 * 1e-3, (rname, np.abs(delta - delta_ref).max())


@pytest.mark.parametrize("gemm_mode", [nv.GEMM_FP32, nv.GEMM_3XTF32], ids=["fp32", "3xtf32"])
def test_loss_curve_matches_reference_1000_steps(gemm_mode):
    """North-star bar: the loss curve of the sm_100a path stays within 1e-3 of the REFERENCE's (float32 release
    build, its own host sampler, cuBLAS SGEMM + cuDNN batch-norm) over 1000 steps of the NVSM recipe
    (batch-norm + hard_tanh, full Adam, lambda 0.01, lr 1e-3) from the same seed."""
    c = dict(V=5000, D=20000, dw=64, dd=64, n=6, z=5, B=2048, nonlinearity=nv.HARD_TANH, bn=True, clip=True,
             bias_neg=False)
    method = (nv.ADAM, nv.DENSE_UPDATE_DENSE_VARIANCE)
    rm = ref_model(c, method, np.float32)
    gm = cuda_model(c, method, gemm_mode=gemm_mode)
    rng = nv.RNG(7)
    gm.initialize(rng)
    nrng = np.random.default_rng(0)
    batches = []
    for _ in range(40):
        f, fw, labels, w = make_batch(nrng, c["B"], c["n"], c["V"], c["D"], c["z"], weighted=False)
        labels = (f[:, 0] * 4 + f[:, 1] % 4) % c["D"]      # learnable: the document depends on the first two words
        batches.append((rm.new_batch().fill(f, labels, fw, w), nv.Batch(c["B"], c["n"]).fill(f, labels, fw, w)))
    lr, g_costs, r_costs = 1e-3, [], []
    for step in range(1000):
        rb, gb = batches[step % len(batches)]
        r_costs.append(rm.step(rb, lr))
        res = gm.compute_cost(gb, rng)
        gm.backprop(res, lr)
        g_costs.append(res.get_cost())
    assert rng.state == rm.rng_state, "both samplers consumed the engine identically for 1000 steps"
    g, r = np.array(g_costs), np.array(r_costs)
    assert r[-1] < r[0] - 0.05, "the run must actually learn"
    dev = np.abs(g - r)
    print("gemm_mode %d: reference loss %.4f -> %.4f, max |cuda - reference| = %.2e (first 100: %.2e, 500: %.2e)" % (
        gemm_mode, r[0], r[-1], dev.max(), dev[:100].max(), dev[:500].max()))
    assert dev.max() <= 1e-3, dev.max()


# ---- 8f-1: RepresentationSimilarity objectives and the mixtures against the reference --------------------------------
SIM_METHODS = [(nv.SGD, 0), (nv.ADAGRAD, 0), (nv.ADAM, nv.SPARSE), (nv.ADAM, nv.DENSE_UPDATE),
               (nv.ADAM, nv.DENSE_UPDATE_DENSE_VARIANCE)]


def _objective_models(objective, c, method, N, wt, ws, lam=0.01):
    rm = R.ObjectiveModel(objective, c["V"], c["D"], c["dw"], c["dd"], batch_size=c["B"], window_size=c["n"],
                          num_random_entities=c["z"], similarity_batch_size=N, nonlinearity=c["nonlinearity"],
                          batch_normalization=c["bn"], clip_sigmoid=c["clip"], bias_negative_samples=c["bias_neg"],
                          update_method=method[0], adam_mode=method[1], regularization_lambda=lam,
                          text_entity_weight=wt, similarity_weight=ws, seed=7, dtype=np.float32)
    desc = nv.ModelDesc(word_repr_size=c["dw"], entity_repr_size=c["dd"], batch_normalization=c["bn"],
                        nonlinearity=c["nonlinearity"], clip_sigmoid=c["clip"], bias_negative_samples=c["bias_neg"])
    ee = objective in (nv.ENTITY_ENTITY, nv.TEXT_ENTITY_ENTITY_ENTITY)
    tc = nv.TrainConfig(batch_size=c["B"], window_size=c["n"], num_random_entities=c["z"], regularization_lambda=lam,
                        update_method=method[0], adam_mode=method[1], text_entity_weight=wt,
                        entity_entity_weight=ws if ee else 0.0, term_term_weight=0.0 if ee else ws)
    gm = nv.Model(c["V"], c["D"], desc, tc, objective=objective, max_similarity_batch_size=N)
    rng = nv.RNG(7)
    gm.initialize(rng)
    assert rng.state == rm.rng_state
    for rname, _, gname in PARAMS:
        np.testing.assert_array_equal(rm.get(rname), gm.get_tensor(gname), err_msg="init " + rname)
    return rm, gm, rng


def _check_params(gm, rm, method, lr, what):
    for rname, _, gname in PARAMS:
        ref_value = rm.get(rname)
        floor = 1e-5 if method[0] != nv.ADAM else 2e-3 * lr / max(np.abs(ref_value).max(), 1e-30)
        assert_close(gm.get_tensor(gname), ref_value, 5e-4, max(floor, 1e-5), "%s %s" % (rname, what))
        gm.set_tensor(gname, ref_value)


@pytest.mark.parametrize("objective", [nv.ENTITY_ENTITY, nv.TERM_TERM], ids=["entity_entity", "term_term"])
@pytest.mark.parametrize("method", SIM_METHODS, ids=lambda m: "m%d_%d" % m)
def test_similarity_objective_matches_reference(objective, method):
    """Model<EntityEntity::Objective> / Model<TermTerm::Objective> (cpp/objective.cu:485-700): cost, pair
    probabilities through the gradient columns, and the one updated table for three steps; the other parameters must
    stay untouched (the reference skips gradient-less parameters, cpp/params.cu:301-304)."""
    c = CASES[1]
    N = 384
    rm, gm, _ = _objective_models(objective, c, method, N, 1.0, 1.0)
    nrng = np.random.default_rng(3)
    limit = c["D"] if objective == nv.ENTITY_ENTITY else c["V"]
    gname = "grad_entity_repr" if objective == nv.ENTITY_ENTITY else "grad_phrase_reprs"
    lr = 0.05
    for step in range(3):
        pairs = nrng.integers(0, limit, size=(N, 2), dtype=np.int64)
        w = nrng.uniform(0.0, 2.0, size=N).astype(np.float32)
        table_name = nv.ENTITY_REPRS if objective == nv.ENTITY_ENTITY else nv.WORD_REPRS
        table = gm.get_tensor(table_name).reshape(limit, -1).copy()      # before the update: the oracle's input
        rm.fill_pairs(pairs, w); rm.forward()
        res = gm.similarity_compute_cost(nv.SimilarityBatch(N).fill(pairs, w))
        rcost = rm.get_cost()
        assert abs(res.get_cost() - rcost) <= 2e-4 * abs(rcost) + 1e-7
        assert abs(res.scaled_regularization_lambda() - rm.scaled_lambda()) <= 1e-9
        rm.compute_gradients(); gm.compute_gradients(res)
        assert_close(gm.get_tensor("grad_similarity"), rm.get(gname), 2e-4, 1e-5, "pair gradient step %d" % step)
        # third witness: the CPU restatement (oracle/nvsm_oracle.hpp: similarity_step)
        ocost, _, ograd = O.similarity_step(table, pairs, w, clip_sigmoid=c["clip"], dtype=np.float32)
        assert abs(res.get_cost() - ocost) <= 2e-4 * abs(ocost) + 1e-7
        assert_close(gm.get_tensor("grad_similarity"), ograd, 2e-4, 1e-5, "pair gradient vs oracle, step %d" % step)
        rm.update(lr, rm.scaled_lambda()); gm.update(None, lr, res.scaled_regularization_lambda())
        _check_params(gm, rm, method, lr, "after step %d" % step)


@pytest.mark.parametrize("objective", [nv.TEXT_ENTITY_ENTITY_ENTITY, nv.TEXT_ENTITY_TERM_TERM], ids=["te_ee", "te_tt"])
@pytest.mark.parametrize("method", [(nv.SGD, 0), (nv.ADAM, nv.DENSE_UPDATE), (nv.ADAM, nv.DENSE_UPDATE_DENSE_VARIANCE)],
                         ids=lambda m: "m%d_%d" % m)
@pytest.mark.parametrize("case", [1, 2])
def test_mixture_objective_matches_reference(objective, method, case):
    """TextEntityEntityEntity / TextEntityTermTerm (cpp/objective.cu:700-794): merged gradients with weights
    w_k / sum w_k, averaged cost and lambda, every optimiser that supports multiple gradient descriptors."""
    c = CASES[case]
    N = 256
    wt, ws = 0.7, 0.3
    rm, gm, rng = _objective_models(objective, c, method, N, wt, ws)
    nrng = np.random.default_rng(9)
    limit = c["D"] if objective == nv.TEXT_ENTITY_ENTITY_ENTITY else c["V"]
    lr = 0.05
    for step in range(3):
        f, fw, labels, w = make_batch(nrng, c["B"], c["n"], c["V"], c["D"], c["z"])
        pairs = nrng.integers(0, limit, size=(N, 2), dtype=np.int64)
        pw = nrng.uniform(0.0, 2.0, size=N).astype(np.float32)
        rm.fill_text(f, labels, fw, w); rm.fill_pairs(pairs, pw); rm.forward()
        res = gm.compute_cost_mixture(nv.Batch(c["B"], c["n"]).fill(f, labels, fw, w), nv.SimilarityBatch(N).fill(pairs, pw), rng)
        assert rng.state == rm.rng_state
        rcost = rm.get_cost()
        assert abs(res.get_cost() - rcost) <= 2e-4 * abs(rcost) + 1e-7
        assert abs(res.scaled_regularization_lambda() - rm.scaled_lambda()) <= 1e-8
        rm.compute_gradients(); gm.compute_gradients(res)
        assert_close(gm.get_tensor("grad_transform"), rm.get("grad_transform"), 2e-4, 1e-5, "merged gT step %d" % step)
        assert_close(gm.get_tensor("grad_bias"), rm.get("grad_bias"), 2e-4, 1e-5, "merged gb step %d" % step)
        rm.update(lr, rm.scaled_lambda()); gm.update(None, lr, res.scaled_regularization_lambda())
        _check_params(gm, rm, method, lr, "after step %d" % step)


def test_mixture_rejects_single_descriptor_optimisers():
    """Adagrad and sparse Adam abort on multiple gradient descriptors in the reference
    (cpp/updates_adagrad.cu:108-109, cpp/updates_adam.cu:339-340); the library returns an error instead."""
    c = CASES[0]
    desc = nv.ModelDesc(word_repr_size=c["dw"], entity_repr_size=c["dd"], clip_sigmoid=True)
    tc = nv.TrainConfig(batch_size=c["B"], window_size=c["n"], num_random_entities=c["z"], update_method=nv.ADAGRAD,
                        text_entity_weight=0.5, entity_entity_weight=0.5)
    gm = nv.Model(c["V"], c["D"], desc, tc, objective=nv.TEXT_ENTITY_ENTITY_ENTITY, max_similarity_batch_size=16)
    rng = nv.RNG(7)
    gm.initialize(rng)
    f, fw, labels, w = make_batch(np.random.default_rng(0), c["B"], c["n"], c["V"], c["D"], c["z"])
    res = gm.compute_cost_mixture(nv.Batch(c["B"], c["n"]).fill(f, labels, fw, w),
                                  nv.SimilarityBatch(16).fill(np.zeros((16, 2), np.int64) + [[1, 2]]), rng)
    gm.compute_gradients(res)
    with pytest.raises(nv.NvsmError):
        gm.update(None, 0.01, res.scaled_regularization_lambda())
