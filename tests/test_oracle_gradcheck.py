"""Central-difference gradient check of the oracle itself (float64), the method of the reference's
cpp/gradient_check.cu:3-133 / gradient_checking_tests.cu: for every parameter tensor, perturb
entries by +-eps, recompute the cost with the same sampled ids, and compare with the analytic
gradient. Gradients are ascent directions (cpp/objective.cu:324-326), hence the sign flip
(gradient_check.cu:43). Covers tanh / hard_tanh x batch-norm on/off x bias_negative_samples."""
import itertools

import numpy as np
import pytest

from oracle import binding as O


def dense_table_grad(sparse_grad, ids, window, weights, num_objects, dim):
    g = np.zeros((num_objects, dim))
    sg = sparse_grad.reshape(-1, dim)
    ids = ids.reshape(-1, window)
    w = np.ones_like(ids, dtype=np.float64) if weights is None else weights.reshape(-1, window)
    for x in range(ids.shape[0]):
        for y in range(window):
            g[ids[x, y]] += w[x, y] * sg[x]
    return g


@pytest.mark.parametrize("nonlin,bn,bias_neg", list(itertools.product([O.TANH, O.HARD_TANH], [False, True], [False, True])))
def test_oracle_gradients_match_central_differences(nonlin, bn, bias_neg):
    V, D, dw, dd, n, z, B = 20, 15, 3, 4, 3, 2, 64   # the reference's check uses V=20 d_w=3 D=15 d_d=4
    rng = np.random.default_rng(11)
    m = O.Model(V, D, dw, dd, nonlinearity=nonlin, batch_normalization=bn, clip_sigmoid=False,
                bias_negative_samples=bias_neg, num_random_entities=z, regularization_lambda=0.01)
    m.initialize(7)
    # spread the pre-activations so hard_tanh has both clipped and unclipped units
    m.set("T", rng.normal(0, 1.0, dw * dd)); m.set("b", rng.normal(0, 0.3, dd))
    f = rng.integers(0, V, (B, n)); fw = rng.uniform(0.5, 1.5, (B, n))
    labels = rng.integers(0, D, B); w = rng.uniform(0.5, 1.5, B)
    ids, _ = O.generate_labels(labels, z, D, 3)

    def cost():
        return m.compute_cost(f, fw, ids, w, n)

    cost()
    m.compute_gradients()
    analytic = {
        "T": m.get("gT"), "b": m.get("gb"),
        "W": dense_table_grad(m.get("gP"), f, n, fw, V, dw).ravel(),
        "E": dense_table_grad(m.get("gE"), ids, 1, None, D, dd).ravel(),
    }
    eps = 1e-6
    for name, grad in analytic.items():
        theta = m.get(name)
        idxs = rng.choice(theta.size, size=min(theta.size, 25), replace=False)
        for i in idxs:
            t = theta.copy(); t[i] += eps; m.set(name, t); cp = cost()
            t[i] -= 2 * eps; m.set(name, t); cm = cost()
            m.set(name, theta)
            numeric = -(cp - cm) / (2 * eps)        # ascent direction
            assert abs(numeric - grad[i]) <= 1e-5 * max(1.0, abs(grad[i])) + 2e-7, (name, i, numeric, grad[i])


@pytest.mark.parametrize("clip", [False, True])
def test_similarity_objective_gradients_match_central_differences(clip):
    """RepresentationSimilarity objective (EntityEntity / TermTerm, cpp/objective.cu:485-672): closed form of the cost,
    and the analytic gradient (ascent direction, a member's gradient is its partner's row) against central differences."""
    rng = np.random.default_rng(7)
    rows, dim, N = 9, 5, 40
    table = rng.normal(scale=0.6, size=(rows, dim))
    ids = rng.integers(0, rows, size=(N, 2))
    ids[3] = (4, 4)                                  # a pair of an object with itself: both halves land on one row
    w = rng.uniform(0.5, 2.0, size=N)
    cost, probs, grad = O.similarity_step(table, ids, w, clip_sigmoid=clip)
    s = (table[ids[:, 0]] * table[ids[:, 1]]).sum(axis=1)
    p = 1.0 / (1.0 + np.exp(-s))
    np.testing.assert_allclose(probs, p, rtol=1e-12)
    np.testing.assert_allclose(cost, -(w * np.log(p)).sum() / N, rtol=1e-12)
    dense = dense_table_grad(grad.ravel(), ids.ravel(), 1, None, rows, dim)
    eps = 1e-6
    for r in range(rows):
        for k in range(dim):
            t = table.copy(); t[r, k] += eps
            up = O.similarity_step(t, ids, w, clip_sigmoid=clip)[0]
            t[r, k] -= 2 * eps
            down = O.similarity_step(t, ids, w, clip_sigmoid=clip)[0]
            numeric = (up - down) / (2 * eps)
            assert abs(-dense[r, k] - numeric) <= 1e-6 * max(1.0, abs(numeric)), (r, k, dense[r, k], numeric)


def test_similarity_objective_saturation_band():
    """clip_sigmoid: probabilities clamp at 1e-7 / 1 - 1e-7 and the multiplier vanishes outside (1e-6, 1 - 1e-6)
    (cpp/objective.cu:545-549, 612-616), exactly like the TextEntity objective."""
    table = np.array([[30.0, 0.0], [1.0, 0.0], [-1.0, 0.0], [0.1, 0.2]])
    ids = np.array([[0, 1], [0, 2], [3, 3]])
    cost, probs, grad = O.similarity_step(table, ids, np.ones(3), clip_sigmoid=True)
    assert probs[0] == 1.0 - 1e-7 and probs[1] == 1e-7
    assert (grad[0:4] == 0).all() and (grad[4:6] != 0).any()
    _, probs_open, grad_open = O.similarity_step(table, ids, np.ones(3), clip_sigmoid=False)
    assert probs_open[1] < 1e-12 and (grad_open[2:4] != 0).any()
