"""Pin the CPU oracle against the reference's own known-answer tests (float64,
like every reference test: cpp/CMakeLists.txt:18). Values live in
tests/golden/reference_goldens.json (extracted by extract_reference_goldens.py)
or are the closed forms the reference tests state."""
import itertools
import json
import math
import os

import numpy as np
import pytest

from oracle import binding as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_goldens.json")))
F = np.float64
RTOL = 1e-12  # the reference uses 4-ULP gtest matchers; summation order inside its
              # un-vendored device_matrix / cuBLAS is unknown, so allow a few more ULPs.


def close(a, b, rtol=RTOL, atol=0.0):
    np.testing.assert_allclose(np.asarray(a, dtype=F), np.asarray(b, dtype=F), rtol=rtol, atol=atol)


# --- cpp/model_tests.cu:52-123 -------------------------------------------------
def test_get_average_representations():
    reprs = np.arange(12, dtype=F)  # 4 objects x 3 dims, counting init (:17-31)
    out = O.gather_mean(reprs, 3, [1, 3, 2, 0, 3, 1], None, 3)
    close(out.ravel(), [(3 + 9 + 6) / 3., (4 + 10 + 7) / 3., (5 + 11 + 8) / 3.,
                        (0 + 9 + 3) / 3., (1 + 10 + 4) / 3., (2 + 11 + 5) / 3.])


def test_get_weighted_average_representations():
    reprs = np.arange(12, dtype=F)
    w = [0.5, 0.3, 0.1, 1.0, 2.0, 0.2]
    out = O.gather_mean(reprs, 3, [1, 3, 2, 0, 3, 1], w, 3)
    close(out.ravel(), [(0.5 * 3 + 0.3 * 9 + 0.1 * 6) / 3., (0.5 * 4 + 0.3 * 10 + 0.1 * 7) / 3.,
                        (0.5 * 5 + 0.3 * 11 + 0.1 * 8) / 3., (1.0 * 0 + 2.0 * 9 + 0.2 * 3) / 3.,
                        (1.0 * 1 + 2.0 * 10 + 0.2 * 4) / 3., (1.0 * 2 + 2.0 * 11 + 0.2 * 5) / 3.])


# --- cpp/model_tests.cu:125-151 --------------------------------------------------
@pytest.mark.parametrize("seed", range(1, 11))
def test_generate_labels(seed):
    ids, _ = O.generate_labels([1, 2, 3, 4, 5], 10, 5000, seed)
    assert ids.size == 5 * 11
    assert list(ids[::11]) == [1, 2, 3, 4, 5]
    assert ids.min() >= 0 and ids.max() < 5000


# --- cpp/model_tests.cu:153-243 --------------------------------------------------
def test_representations_update_decay_only():
    table = np.arange(12, dtype=F)
    up = O.ReprUpdater(O.SGD, 0, 4, 3)
    g = np.zeros((2, 3), dtype=F)
    lam = 0.1 / 1  # batch_size_ of that ForwardResult = |entities| / (z + 1)... see below
    # ForwardResult(words={0,3,1,0}, entities={0}, window 2, z=1, lambda=0.1): batch_size = 4/2 = 2
    # => scaled lambda = 0.1 / 2 (the test's scale_factor = 1 - (0.1 * 0.1) / 2).
    up.update(table, [(g, [0, 3, 1, 0], 2, np.ones(4))], 0.1, 0.1 / 2.0)
    close(table, np.arange(12) * (1.0 - (0.1 * 0.1) / 2.0))


def test_representations_update_scatter():
    table = np.arange(12, dtype=F)
    up = O.ReprUpdater(O.SGD, 0, 4, 3)
    g = np.array([[5.0, 4.0, 3.0], [-3.0, -2.0, 10.0]], dtype=F)
    up.update(table, [(g, [0, 3, 1, 0], 2, np.ones(4))], 0.1, 0.0)
    lr = 0.1
    close(table, [0. + (5.0 + (-3.0)) * lr, 1. + (4.0 + (-2.0)) * lr, 2. + (3.0 + 10.0) * lr,
                  3. + (-3.0) * lr, 4. + (-2.0) * lr, 5. + 10.0 * lr, 6., 7., 8.,
                  9. + 5.0 * lr, 10. + 4.0 * lr, 11. + 3.0 * lr])


# --- cpp/model_tests.cu:245-275 --------------------------------------------------
def test_update_dense():
    table = np.arange(12, dtype=F)
    O.update_dense(table, np.full(12, 10.0), 0.1, 0.01)
    close(table, [i * (1.0 - 0.01 * 0.1) + 10.0 * 0.1 for i in range(12)])


# --- cpp/model_tests.cu:277-339 --------------------------------------------------
def _transform_model(bn, eps=1e-4):
    m = O.Model(1, 1, 3, 5, nonlinearity=O.TANH, batch_normalization=bn, bn_epsilon=eps)
    m.set("T", np.arange(15, dtype=F))          # counting init, column-major 5 x 3
    m.set("b", np.arange(5, dtype=F) * 1e-3)
    return m


def test_transform_forward():
    m = _transform_model(False)
    m.set("W", [0, 0, 0])
    # Drive Transform::transform through infer-like path: P supplied via a 2-word table.
    m2 = O.Model(2, 1, 3, 5, nonlinearity=O.TANH)
    m2.set("T", np.arange(15, dtype=F)); m2.set("b", np.arange(5, dtype=F) * 1e-3)
    m2.set("W", [0.01, 0.02, 0.03, 0.001, 0.002, 0.003])
    out = m2.infer([0, 1], 1)
    close(out.ravel(), np.tanh([0.400, 0.461, 0.522, 0.583, 0.644, 0.040, 0.047, 0.054, 0.061, 0.068]))


# --- cpp/model_tests.cu:468-521 (BN eps 1e-5 in that test) ------------------------
def test_transform_batch_normalization_forward():
    g = GOLD["transform_bn_forward"]
    m = O.Model(2, 1, 3, 5, nonlinearity=O.TANH, batch_normalization=True, bn_epsilon=g["epsilon"],
                num_random_entities=1)
    m.set("T", np.arange(15, dtype=F)); m.set("b", np.arange(5, dtype=F) * 1e-3)
    m.set("W", [0.01, 0.02, 0.03, 0.001, 0.002, 0.003])
    m.set("E", np.zeros(5))
    m.compute_cost([0, 1], [1.0, 1.0], [0, 0, 0, 0], [1.0, 1.0], 1)
    close(m.get("Y"), g["output"])
    # backward stays finite (:523-547)
    m.compute_gradients()
    assert np.isfinite(m.get("gT")).all() and np.isfinite(m.get("gb")).all()


# --- cpp/model_tests.cu:341-466: the anchor known-answer test ----------------------
def test_transform_backward_full_chain():
    g = GOLD["transform_backward"]
    c = g["config"]
    m = O.Model(c["num_words"], c["num_entities"], c["word_repr_size"], c["entity_repr_size"],
                nonlinearity=O.TANH, bias_negative_samples=c["bias_negative_samples"],
                num_random_entities=c["num_random_entities"],
                regularization_lambda=c["regularization_lambda"], update_method=O.SGD)
    state = m.initialize(c["seed"])
    B, n, z = c["batch_size"], c["window_size"], c["num_random_entities"]
    labels = np.full(B, c["label"], dtype=np.int64)
    ids, state = O.generate_labels(labels, z, c["num_entities"], state)
    cost = m.compute_cost(np.full(B * n, c["feature_value"]), np.ones(B * n), ids, np.ones(B), n)
    assert math.isfinite(cost)
    m.compute_gradients()
    close(m.get("gT"), g["grad_transform"])
    close(m.get("gb"), g["grad_bias"])
    close(m.get("gP"), g["grad_phrase"])
    assert abs(cost - 6.17158013374) < 1e-9  # value printed by the survey's probe


# --- cpp/cudnn_utils_tests.cu -------------------------------------------------------
def test_bn_constant_input_is_zero():  # :19-36
    bn = O.BatchNorm(1e-4)
    y = bn.forward(np.ones((100, 10)), np.zeros(10))
    assert (y == 0).all()


def test_bn_forward_backward_golden():  # :115-177
    g = GOLD["bn_forward_backward"]
    eps = g["epsilon"]
    bn = O.BatchNorm(eps)
    x = np.array(g["input"], dtype=F).reshape(2, 3)
    y = bn.forward(x, np.zeros(3))
    close(y.ravel(), [(1.0 - 3.0) / math.sqrt(4.0 + eps), (2.0 - 6.0) / math.sqrt(16.0 + eps),
                      (3.0 - 11.5) / math.sqrt(72.25 + eps), (5.0 - 3.0) / math.sqrt(4.0 + eps),
                      (10.0 - 6.0) / math.sqrt(16.0 + eps), (20.0 - 11.5) / math.sqrt(72.25 + eps)])
    dx, db = bn.backward(np.array(g["grad_output"], dtype=F).reshape(2, 3), x)
    close(db, g["grad_bias"])
    # dx is a difference of nearly equal terms (~1e-7 out of ~1): absolute tolerance.
    close(dx.ravel(), g["grad_input"], rtol=0, atol=1e-15)


# --- cpp/cuda_utils_tests.cu:8-21 -----------------------------------------------------
def test_truncated_sigmoid():
    s = lambda x, e: O.scalar_fn("truncated_sigmoid", F, x, e)
    assert s(0.0, 0.0) == 0.5
    close(s(1.0, 0.0), 0.7310585786300049)
    close(s(-1.0, 0.0), 1.0 - 0.7310585786300049)
    assert s(-50.0, 0.0) > 0.0 and s(20.0, 0.0) < 1.0
    close(s(-100.0, 1e-7), 1e-7)
    close(s(100.0, 1e-7), 1.0 - 1e-7)


def test_hard_tanh_bounds():  # include/cuNVSM/cuda_utils.h:86-147
    clip = lambda x: O.scalar_fn("clip", F, x)
    d = lambda y: O.scalar_fn("clip_deriv", F, y)
    assert clip(5.0) == np.nextafter(1.0, 2.0) and clip(-5.0) == np.nextafter(-1.0, -2.0)
    assert d(1.0) == 1.0 and d(-1.0) == 1.0 and d(clip(5.0)) == 0.0 and d(clip(-5.0)) == 0.0


# --- cpp/updates_tests.cu ----------------------------------------------------------------
PARAMS = list(itertools.product([0.0, 0.1], [1.0, 0.5]))  # (scaled lambda, lr) :28-32
G24 = np.arange(1.0, 25.0)
GB = np.array([25.0, 26.0, 27.0])


@pytest.mark.parametrize("lam,lr", PARAMS)
def test_sgd_transform(lam, lr):  # :34-97
    T = np.full(24, 5.0); b = np.full(3, 5.0)
    up = O.TransformUpdater(O.SGD, 24, 3)
    up.update(T, b, G24.copy(), GB.copy(), lr, lam)
    close(T, 5.0 + lr * (G24 - lam * 5.0))
    close(b, 5.0 + lr * GB)


def _two_descs():
    g1 = np.array([[2.0, 2.5, 3.0, 4.0]]); g2 = np.array([[10.0, 11.0, 12.0, 13.0]])
    return g1, g2


@pytest.mark.parametrize("lam,lr", PARAMS)
def test_sgd_representations(lam, lr):  # :99-172
    table = np.full(40, 5.0)
    up = O.ReprUpdater(O.SGD, 0, 10, 4)
    g1, g2 = _two_descs()
    up.update(table, [(g1, [9, 0, 1], 3, None), (g2, [5, 1, 8], 3, None)], lr, lam)
    exp = np.full((10, 4), (1.0 - lr * lam) * 5.0)
    exp[0] = 5.0 + lr * (g1[0] - lam * 5.0)
    exp[1] = 5.0 + lr * (g1[0] + g2[0] - lam * 5.0)
    exp[5] = 5.0 + lr * (g2[0] - lam * 5.0)
    exp[8] = 5.0 + lr * (g2[0] - lam * 5.0)
    exp[9] = 5.0 + lr * (g1[0] - lam * 5.0)
    close(table, exp.ravel())


@pytest.mark.parametrize("lam,lr", PARAMS)
def test_adagrad_transform(lam, lr):  # :174-248
    eps = 1e-6
    T = np.full(24, 5.0); b = np.full(3, 5.0)
    up = O.TransformUpdater(O.ADAGRAD, 24, 3, eps=eps)
    gT, gb = G24.copy(), GB.copy()
    up.update(T, b, gT, gb, lr, lam)
    close(up.state(0), G24 ** 2)
    close(up.state(1), GB ** 2)
    close(gT, G24 / np.sqrt(G24 ** 2 + eps))
    close(gb, GB / np.sqrt(GB ** 2 + eps))


@pytest.mark.parametrize("lam,lr", PARAMS)
def test_adagrad_representations(lam, lr):  # :250-297
    eps = 1e-6
    table = np.full(40, 5.0)
    up = O.ReprUpdater(O.ADAGRAD, 0, 10, 4, eps=eps)
    g = np.array([[2.0, 2.5, 3.0, 4.0], [10.0, 11.0, 12.0, 13.0]])
    g0 = g.copy()
    up.update(table, [(g, [9, 0, 1, 5, 1, 8], 3, None)], lr, lam)
    close(up.state(0), [8.8125, 142.3125, 0.0, 0.0, 0.0, 133.5, 0.0, 0.0, 133.5, 8.8125])
    close(g[0], g0[0] / math.sqrt(((8.8125 + 8.8125 + 142.3125) / 3.0) + eps))
    close(g[1], g0[1] / math.sqrt(((133.5 + 142.3125 + 133.5) / 3.0) + eps))


@pytest.mark.parametrize("lam,lr", PARAMS)
def test_adam_transform_two_steps(lam, lr):  # :299-425
    gold = GOLD["adam_transform"]
    eps, b1, b2 = gold["epsilon"], 0.9, 0.999
    T = np.full(24, 5.0); b = np.full(3, 5.0)
    up = O.TransformUpdater(O.ADAM, 24, 3, beta1=b1, beta2=b2, eps=eps)
    gT, gb = G24.copy(), GB.copy()
    up.update(T, b, gT, gb, lr, lam)
    bc1 = math.sqrt(1.0 - b2) / (1.0 - b1)
    g = G24 - lam * 5.0
    close(gT, bc1 * ((1.0 - b1) * g) / (np.sqrt((1.0 - b2) * g ** 2) + eps))
    close(gb, gold["grad_bias_t1"])
    close(up.state(1), gold["m_bias_t1"])
    close(up.state(3), gold["v_bias_t1"])
    T_before = T.copy()
    gT, gb = G24.copy(), GB.copy()
    up.update(T, b, gT, gb, lr, lam)
    bc2 = math.sqrt(1.0 - b2 ** 2) / (1.0 - b1 ** 2)
    g2 = G24 - lam * T_before
    m1 = (1.0 - b1) * g; v1 = (1.0 - b2) * g ** 2
    close(gT, bc2 * (b1 * m1 + (1.0 - b1) * g2) / (np.sqrt(b2 * v1 + (1.0 - b2) * g2 ** 2) + eps))
    close(gb, gold["grad_bias_t2"])       # non-decaying bias moments
    close(up.state(1), gold["m_bias_t2"])
    close(up.state(3), gold["v_bias_t2"])


def _adam_expected_m(b1, lam_term=0.0):
    g1 = np.array([2.0, 2.5, 3.0, 4.0]); g2 = np.array([10.0, 11.0, 12.0, 13.0])
    rows = [g1, g1 + g2, g2, g2, g1]
    return np.concatenate([(1.0 - b1) * (r - lam_term) for r in rows])


@pytest.mark.parametrize("lam,lr", PARAMS)
def test_adam_representations_sparse(lam, lr):  # :427-525
    eps, b1, b2 = 1e-5, 0.9, 0.999
    table = np.full(20, 5.0)
    up = O.ReprUpdater(O.ADAM, O.SPARSE, 5, 4, beta1=b1, beta2=b2, eps=eps)
    g = np.array([[2.0, 2.5, 3.0, 4.0], [10.0, 11.0, 12.0, 13.0]])
    g0 = g.copy()
    up.update(table, [(g, [4, 0, 1, 3, 1, 2], 3, None)], lr, lam)
    close(up.state(1), _adam_expected_m(b1))
    vexp = (1.0 - b2) * np.array([8.8125, 8.8125 + 133.5, 133.5, 133.5, 8.8125])
    close(up.state(2), vexp)
    bc = math.sqrt(1.0 - b2) / (1.0 - b1)
    e0 = bc * ((1.0 - b1) * (g0[0] + g0[0] + (g0[0] + g0[1])) / 3) / (math.sqrt((1.0 - b2) * (8.8125 + 8.8125 + 133.5 + 8.8125) / 3) + eps)
    e1 = bc * ((1.0 - b1) * (g0[1] + (g0[0] + g0[1]) + g0[1]) / 3) / (math.sqrt((1.0 - b2) * (133.5 + (8.8125 + 133.5) + 133.5) / 3) + eps)
    close(g[0], e0); close(g[1], e1)


@pytest.mark.parametrize("lam,lr", PARAMS)
def test_adam_representations_dense_update(lam, lr):  # :527-631
    eps, b1, b2 = 1e-5, 0.9, 0.999
    table = np.full(20, 5.0)
    up = O.ReprUpdater(O.ADAM, O.DENSE_UPDATE, 5, 4, beta1=b1, beta2=b2, eps=eps)
    g1, g2 = _two_descs()
    up.update(table, [(g1, [4, 0, 1], 3, None), (g2, [3, 1, 2], 3, None)], lr, lam)
    close(up.state(1), _adam_expected_m(b1))
    vobj = (1.0 - b2) * np.array([8.8125, 8.8125 + 133.5, 133.5, 133.5, 8.8125])
    close(up.state(2), vobj)
    bc = math.sqrt(1.0 - b2) / (1.0 - b1)
    m = _adam_expected_m(b1).reshape(5, 4)
    exp = 5.0 + lr * (bc * m / (np.sqrt(vobj)[:, None] + eps) - lam * 5.0)
    close(table, exp.ravel())


@pytest.mark.parametrize("lam,lr", PARAMS)
def test_adam_representations_dense_variance(lam, lr):  # :633-775
    eps, b1, b2 = 1e-5, 0.9, 0.999
    table = np.full(20, 5.0)
    up = O.ReprUpdater(O.ADAM, O.DENSE_UPDATE_DENSE_VARIANCE, 5, 4, beta1=b1, beta2=b2, eps=eps)
    g1, g2 = _two_descs()
    up.update(table, [(g1, [4, 0, 1], 3, None), (g2, [3, 1, 2], 3, None)], lr, lam)
    m = _adam_expected_m(b1, lam * 5.0)
    close(up.state(1), m)
    gfull = m / (1.0 - b1)
    v = (1.0 - b2) * gfull ** 2
    close(up.state(2), v)
    bc = math.sqrt(1.0 - b2) / (1.0 - b1)
    close(table, 5.0 + lr * (bc * m / (np.sqrt(v) + eps)))


# --- Glorot init + RNG plumbing -----------------------------------------------------------
def test_glorot_range_and_stream():
    a, st = O.glorot(3, 7, 10)
    lim = math.sqrt(6.0 / 10)
    assert (np.abs(a) <= lim).all() and a.std() > 0
    # generate_canonical<double,1> on minstd_rand0 consumes one draw per value.
    x = 10
    for _ in range(21):
        x = (x * 16807) % 2147483647
    assert st == x


def test_normalizer_golden():  # cpp/cuda_utils_tests.cu:51-92
    """L2 Normalizer forward (closed form) and backward (the reference's ten literals), float64 like the reference's
    test build."""
    g = GOLD["normalizer"]
    x = np.array(g["input"], dtype=np.float64)
    y, dx = O.normalizer(x, np.array(g["grad_output"], dtype=np.float64))
    np.testing.assert_allclose(y, x / np.sqrt((x * x).sum(axis=1, keepdims=True)), rtol=1e-15)
    np.testing.assert_allclose(dx.ravel(), g["grad_input"], rtol=1e-12)
    # in place, like the reference's own call (normalizer.forward(input, &input))
    y32, dx32 = O.normalizer(x, np.array(g["grad_output"]), dtype=np.float32)
    np.testing.assert_allclose(dx32.ravel(), g["grad_input"], rtol=2e-3)     # float32: |dy| ~ 1e4 cancels to ~1e2
