"""-m gpu: parity against the REFERENCE ITSELF (oracle/_ref, float32 release build) at the sizes BASELINE.json
states, through the C ABI, in the arithmetic bench.py times (3xTF32 tcgen05 GEMMs, ring score kernel, pull updates):

  * configs[1] C2  NVSM hard_tanh + batch-norm, B=51200, |V|=|D|=50k, 300/256, n=10, z=10, full Adam: 3 steps
  * configs[2] C3  |V|=200k, |D|=500k, z=16, Adagrad: 2 steps
  * configs[4] C5  LSE tanh, bias_negative_samples, |V|=100k, |D|=1M, d=128, z=32, B=4096, SGD: 2 steps (uniform
                   negatives: UniformLabelGenerator is the reference's only generator)
  * the north-star loss-curve bar ON C2 ITSELF: 1000 steps, |cuda - reference| <= 1e-3 at every step

mirroring the reference's loop (cpp/main.cu:405-444: compute_cost, compute_gradients, update, get_cost).

Tolerances. Step tensors (loss, grad_transform, grad_bias, batch-norm statistics): 2e-4 relative (loss) / 2e-4 of the
tensor's largest magnitude. Parameters after an update are checked three ways, then re-aligned to the reference's so
that every step is checked on its own: (i) relative Frobenius error of the step's DELTA, (ii) the largest elementwise
deviation as a fraction of the learning rate, (iii) the fraction of elements outside 5e-4 relative + floor. (ii) and
(iii) are not zero-tolerance at this size for a reason that is arithmetic, not a bug: among 13 M pre-activations a
handful sit within 1e-6 of hard_tanh's clip bound, where the derivative is 0 on one side and 1 on the other, and
Adam's normalised step turns that one flipped gradient element into a visible fraction of lr for the ten word rows
of that n-gram (DESIGN.md section 5 has the measurement). SGD / Adagrad have no such amplifier: elementwise.
"""
import numpy as np
import pytest

import cunvsm_b200 as nv
from oracle import ref_binding as R

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not R.available(np.float32), reason="oracle/_ref not built (needs /root/reference)")]

PARAMS = (("word_representations", nv.WORD_REPRS), ("entity_representations", nv.ENTITY_REPRS),
          ("transform", nv.TRANSFORM), ("bias", nv.BIAS))
C2 = dict(V=50000, D=50000, dw=300, dd=256, n=10, z=10, B=51200, nonlinearity=nv.HARD_TANH, bn=True, bias_neg=False,
          method=(nv.ADAM, nv.DENSE_UPDATE_DENSE_VARIANCE), lr=1e-3)
C3 = dict(V=200000, D=500000, dw=300, dd=256, n=10, z=16, B=51200, nonlinearity=nv.HARD_TANH, bn=True, bias_neg=False,
          method=(nv.ADAGRAD, 0), lr=1e-2)
C5 = dict(V=100000, D=1000000, dw=128, dd=128, n=10, z=32, B=4096, nonlinearity=nv.TANH, bn=False, bias_neg=True,
          method=(nv.SGD, 0), lr=1e-2)


def models(c, gemm_mode=nv.GEMM_3XTF32, lam=0.01, seed=1):
    rm = R.Model(c["V"], c["D"], c["dw"], c["dd"], batch_size=c["B"], window_size=c["n"], num_random_entities=c["z"],
                 nonlinearity=c["nonlinearity"], batch_normalization=c["bn"], clip_sigmoid=True,
                 bias_negative_samples=c["bias_neg"], update_method=c["method"][0], adam_mode=c["method"][1],
                 regularization_lambda=lam, seed=seed, dtype=np.float32)
    desc = nv.ModelDesc(word_repr_size=c["dw"], entity_repr_size=c["dd"], batch_normalization=c["bn"],
                        nonlinearity=c["nonlinearity"], clip_sigmoid=True, bias_negative_samples=c["bias_neg"])
    tc = nv.TrainConfig(batch_size=c["B"], window_size=c["n"], num_random_entities=c["z"], regularization_lambda=lam,
                        update_method=c["method"][0], adam_mode=c["method"][1])
    gm = nv.Model(c["V"], c["D"], desc, tc, gemm_mode=gemm_mode)
    rng = nv.RNG(seed)
    gm.initialize(rng)
    assert rng.state == rm.rng_state, "Glorot init consumed the shared engine identically"
    for rname, gname in PARAMS:
        np.testing.assert_array_equal(rm.get(rname), gm.get_tensor(gname), err_msg="init " + rname)
    return rm, gm, rng


def batch_arrays(nrng, c):
    f = nrng.integers(0, c["V"], size=(c["B"], c["n"]), dtype=np.int64)
    labels = nrng.integers(0, c["D"], size=c["B"], dtype=np.int64)
    return f, labels, np.ones((c["B"], c["n"]), np.float32), np.ones(c["B"], np.float32)


def rel_to_max(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def run_steps(c, steps, check_params):
    rm, gm, rng = models(c)
    nrng = np.random.default_rng(2024)
    lr = c["lr"]
    stats = []
    for step in range(steps):
        f, labels, fw, w = batch_arrays(nrng, c)
        rm.forward(rm.new_batch().fill(f, labels, fw, w))
        res = gm.compute_cost(nv.Batch(c["B"], c["n"]).fill(f, labels, fw, w), rng)
        assert (gm.entity_ids(c["B"]) == rm.entity_ids()).all(), "sampled ids bit-exact at full size"
        assert rng.state == rm.rng_state
        rcost = rm.get_cost()
        assert abs(res.get_cost() - rcost) <= 2e-4 * abs(rcost), (step, res.get_cost(), rcost)
        rm.compute_gradients(); gm.compute_gradients(res)
        for name in ("grad_transform", "grad_bias"):
            err = rel_to_max(gm.get_tensor(name), rm.get(name))
            assert err <= 2e-4, (name, step, err)
        assert abs(res.scaled_regularization_lambda() - rm.scaled_lambda()) <= 1e-12
        before = {gname: gm.get_tensor(gname).astype(np.float64) for _, gname in PARAMS}
        rm.update(lr, rm.scaled_lambda()); gm.update(None, lr, res.scaled_regularization_lambda())
        for rname, gname in PARAMS:
            ref_value = rm.get(rname)
            got = gm.get_tensor(gname).astype(np.float64)
            stats.append(check_params(rname, step, got, ref_value.astype(np.float64), before[gname], lr))
            gm.set_tensor(gname, ref_value)   # re-align the parameters (not the optimiser state): every step on its own
    return stats


def adam_check(rname, step, got, ref, before, lr):
    d_ref, d_got = ref - before, got - before
    frob = np.linalg.norm(d_got - d_ref) / max(np.linalg.norm(d_ref), 1e-30)
    worst = np.abs(got - ref).max() / lr
    outside = float((np.abs(got - ref) > 5e-4 * np.abs(ref) + 2e-3 * lr).mean())
    # measured on B200 (round 2): Frobenius 2e-6 .. 9e-6, no element outside the band
    assert frob <= 2e-4, (rname, step, "relative Frobenius error of the update", frob)
    assert worst <= 0.05, (rname, step, "largest deviation / lr", worst)
    assert outside <= 1e-5, (rname, step, "fraction of elements outside 5e-4 rel + 2e-3 lr", outside)
    return rname, step, frob, worst, outside


def elementwise_check(rname, step, got, ref, before, lr):
    scale = max(np.abs(ref).max(), 1e-30)
    np.testing.assert_allclose(got, ref, rtol=5e-4, atol=1e-5 * scale, err_msg="%s after step %d" % (rname, step))
    return rname, step, np.abs(got - ref).max() / scale


def test_c2_full_size_three_steps_against_reference():
    for row in run_steps(C2, 3, adam_check):
        print("C2 %s step %d: update Frobenius rel err %.2e, max dev %.3f lr, outside-band fraction %.1e" % row)


def test_c3_full_size_adagrad_two_steps_against_reference():
    for row in run_steps(C3, 2, elementwise_check):
        print("C3 %s step %d: max |cuda - reference| / max|reference| = %.2e" % row)


def test_c5_full_size_sgd_two_steps_against_reference():
    for row in run_steps(C5, 2, elementwise_check):
        print("C5 %s step %d: max |cuda - reference| / max|reference| = %.2e" % row)


def test_c2_loss_curve_1000_steps_against_reference():
    """North-star bar at the BASELINE configuration itself: 1000 steps of C2 (batch 51200, full Adam, lambda 0.01,
    lr 1e-3) from the same seed, the reference with its own host sampler, cuBLAS SGEMM and cuDNN batch-norm; the
    sm_100a path with the device-side arithmetic of the bench. The two engines must stay in lock step (bit-exact
    negatives for 1000 x 512000 draws) and the losses within 1e-3 at every step."""
    c = C2
    rm, gm, rng = models(c)
    nrng = np.random.default_rng(0)
    batches = []
    for _ in range(20):
        f, labels, fw, w = batch_arrays(nrng, c)
        labels = (f[:, 0] * 7 + f[:, 1] % 7) % c["D"]      # learnable: the document depends on the first two words
        batches.append((rm.new_batch().fill(f, labels, fw, w), nv.Batch(c["B"], c["n"]).fill(f, labels, fw, w)))
    lr, g_costs, r_costs = c["lr"], [], []
    for step in range(1000):
        rb, gb = batches[step % len(batches)]
        r_costs.append(rm.step(rb, lr))
        res = gm.compute_cost(gb, rng)
        gm.backprop(res, lr)
        g_costs.append(res.get_cost())
    assert rng.state == rm.rng_state, "both samplers consumed the engine identically for 1000 steps"
    g, r = np.array(g_costs), np.array(r_costs)
    dev = np.abs(g - r)
    print("C2 1000 steps: reference loss %.4f -> %.4f, cuda %.4f -> %.4f, max |cuda - reference| = %.2e "
          "(first 100: %.2e, first 500: %.2e)" % (r[0], r[-1], g[0], g[-1], dev.max(), dev[:100].max(), dev[:500].max()))
    assert r[-1] < r[0], "the run must actually learn"
    assert dev.max() <= 1e-3, dev.max()
