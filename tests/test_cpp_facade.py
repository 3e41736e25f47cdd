"""The C++ host side (include/cuNVSM/*.h façade + cpp/main.cpp CLI) over the C ABI."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "cpp")


def _build():
    """make under an inter-process lock: with pytest-xdist another worker may be executing one of the binaries while this
    one relinks it ("Text file busy")."""
    import fcntl
    from cunvsm_b200 import build as b
    with open(os.path.join(CPP, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        b.build()
        subprocess.run(["make", "-C", CPP], check=True, capture_output=True)


def test_library_exports_every_declared_symbol():
    """Every prototype in include/nvsm_b200.h is exported by libnvsm_b200.so and declared in the
    ctypes binding (no compute calls: this runs without a GPU)."""
    import re
    from cunvsm_b200 import _lib
    L = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "nvsm_b200.h")).read()
    declared = set(re.findall(r"NVSM_API [\w\s\*]*?(nvsm_\w+)\(", hdr))
    assert declared, "no prototypes parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.nvsm_version() >= 100
    assert L.nvsm_num_phases() > 5 and L.nvsm_phase_name(1) == b"gather_mean"


def test_host_sampler_is_bit_exact_with_oracle():
    """nvsm_generate_labels (host, no GPU needed) against the oracle's restatement of
    cpp/labels.cu:3-22 for several table sizes, including D where rejection is frequent."""
    from cunvsm_b200 import _lib
    from oracle import binding as O
    L = _lib.load()
    for D, z, B, seed in ((3, 10, 32, 10), (5000, 10, 100, 1), (50000, 10, 2000, 123), (1 << 30, 4, 500, 7),
                          ((1 << 31) - 5, 3, 300, 99), (1, 2, 10, 5)):
        labels = (np.arange(B, dtype=np.int64) * 7) % D
        exp, st_exp = O.generate_labels(labels, z, D, seed)
        out = np.zeros(B * (z + 1), dtype=np.int64)
        st = ctypes.c_ulong(seed)
        pl = ctypes.POINTER(ctypes.c_long)
        assert L.nvsm_generate_labels(labels.ctypes.data_as(pl), B, z, D, ctypes.byref(st), out.ctypes.data_as(pl)) == 0
        assert (out == exp).all() and st.value == st_exp


def test_cpp_host_code_builds_and_fails_loudly_without_gpu():
    _build()
    assert os.path.exists(os.path.join(CPP, "cuNVSMTrainModel")) and os.path.exists(os.path.join(CPP, "facade_test"))
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    res = subprocess.run([os.path.join(CPP, "facade_test")], capture_output=True, text=True)
    assert res.returncode != 0 and "no CUDA device" in res.stderr
    # the Params / Storage / Updates / BatchNormalization class surface: same rule, no silent CPU path
    assert os.path.exists(os.path.join(CPP, "classes_test"))
    res = subprocess.run([os.path.join(CPP, "classes_test")], capture_output=True, text=True)
    assert res.returncode != 0 and "no CUDA device" in res.stderr


@pytest.mark.gpu
def test_reference_class_surface_closed_form_cases():
    """cpp/classes_test.cpp: Representations / Transform / *Storage / the nine *GradientUpdater classes / BatchNormalization
    (include/cuNVSM/{params,storage,updates,cudnn_utils}.h over nvsm_op_* / nvsm_updater_*) on the inputs and closed forms of
    the reference's own unit tests (cpp/updates_tests.cu:34-775, cpp/model_tests.cu:52-339,468-521,
    cpp/cudnn_utils_tests.cu:19-177), all four (lambda, learning rate) parameterisations."""
    _build()
    res = subprocess.run([os.path.join(CPP, "classes_test")], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "CLASSES_TEST_OK" in res.stdout, res.stdout[-4000:] + res.stderr[-2000:]
    assert res.stdout.count("\nok  ") + res.stdout.startswith("ok  ") >= 40, res.stdout[-2000:]


def test_data_source_wrappers():
    """MultiSource / RepeatingSource / LoadSimilarities against the reference's own test cases
    (cpp/data_tests.cpp:828-908), on a plain batch type: no GPU needed."""
    _build()
    res = subprocess.run([os.path.join(CPP, "data_test")], capture_output=True, text=True)
    assert res.returncode == 0 and "data tests ok" in res.stdout, res.stderr


@pytest.mark.gpu
def test_data_sources_on_pinned_batches():
    """RepresentationSimilarity::DataSource (shared-RNG shuffle, partial last batch) and AsyncSource on page-locked batches."""
    _build()
    res = subprocess.run([os.path.join(CPP, "data_test"), "--pinned"], capture_output=True, text=True)
    assert res.returncode == 0 and "pinned batches included" in res.stdout, res.stderr


def test_cli_rejects_bad_flags():
    _build()
    cli = os.path.join(CPP, "cuNVSMTrainModel")
    for args, msg in ((["--update_method", "bogus", "--nonlinearity", "tanh", "--seed", "1"], "valid --update_method"),
                      (["--update_method", "sgd", "--nonlinearity", "relu", "--seed", "1"], "valid --nonlinearity"),
                      (["--update_method", "sgd", "--nonlinearity", "tanh", "--seed", "1", "--weighting", "bogus"], "valid --weighting"),
                      (["--update_method", "sgd", "--nonlinearity", "tanh", "--seed", "1", "--feature_weighting", "tfidf"],
                       "valid --feature_weighting"),
                      (["--update_method", "sgd", "--nonlinearity", "tanh"], "--seed")):
        res = subprocess.run([cli] + args, capture_output=True, text=True)
        assert res.returncode != 0 and msg in res.stderr, res.stderr


@pytest.mark.gpu
def test_cpp_facade_matches_python_driver():
    """Same three full_adam steps through the C++ Model façade and through ctypes: identical costs
    (same library, same RNG stream), identical parameter checksum."""
    _build()
    res = subprocess.run([os.path.join(CPP, "facade_test")], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    lines = dict((" ".join(l.split()[:-1]), l.split()[-1]) for l in res.stdout.strip().splitlines())
    import cunvsm_b200 as nv
    desc = nv.ModelDesc(word_repr_size=16, entity_repr_size=8, batch_normalization=True, nonlinearity=nv.HARD_TANH, clip_sigmoid=True)
    tc = nv.TrainConfig(batch_size=256, window_size=4, num_random_entities=3, regularization_lambda=0.01,
                        update_method=nv.ADAM, adam_mode=nv.DENSE_UPDATE_DENSE_VARIANCE)
    m = nv.Model(100, 60, desc, tc, gemm_mode=nv.GEMM_FP32)
    rng = nv.RNG(5)
    m.initialize(rng)
    i = np.arange(256)
    f = np.stack([i % 100, (i * 7) % 100, (i * 13 + 1) % 100, (i + 50) % 100], 1)
    batch = nv.Batch(256, 4).fill(f, i % 60)
    for step in range(3):
        r = m.compute_cost(batch, rng)
        m.compute_gradients(r)
        m.update(None, 0.001, r.scaled_regularization_lambda())
        assert abs(r.get_cost() - float(lines["cost %d" % step])) <= 2e-6 * abs(r.get_cost())
    assert rng.state == int(lines["rng"])
    cs = sum(float(v.astype(np.float64).sum()) for v in m.get_data().values())
    assert abs(cs - float(lines["checksum"])) <= 1e-4 * abs(cs) + 1e-3
    assert int(lines["params"]) == m.num_parameters()
    # the TextEntityEntityEntity mixture through Model<TextEntityEntityEntity::Objective>
    mtc = nv.TrainConfig(batch_size=256, window_size=4, num_random_entities=3, regularization_lambda=0.01,
                         update_method=nv.ADAM, adam_mode=nv.DENSE_UPDATE_DENSE_VARIANCE, text_entity_weight=0.75,
                         entity_entity_weight=0.25)
    mm = nv.Model(100, 60, desc, mtc, gemm_mode=nv.GEMM_FP32, objective=nv.TEXT_ENTITY_ENTITY_ENTITY)
    mrng = nv.RNG(5)
    mm.initialize(mrng)
    pairs = nv.SimilarityBatch(256).fill(np.stack([i % 60, (i * 11 + 3) % 60], 1), 1.0 + (i % 3))
    for step in range(3):
        r = mm.compute_cost_mixture(batch, pairs, mrng)
        mm.backprop(r, 0.001)
        assert abs(r.get_cost() - float(lines["mixcost %d" % step])) <= 2e-6 * abs(r.get_cost())
    mcs = sum(float(v.astype(np.float64).sum()) for v in mm.get_data().values())
    assert abs(mcs - float(lines["mixchecksum"])) <= 1e-4 * abs(mcs) + 1e-3
    # a mixture result read one step late (the CLI's deferred loss read) reports its own batch's cost
    assert lines["deferred_mixture_cost_ok"] == "1"


@pytest.mark.gpu
def test_cli_trains_on_synthetic_source(tmp_path):
    _build()
    out = str(tmp_path / "model")
    res = subprocess.run([os.path.join(CPP, "cuNVSMTrainModel"), "--num_epochs", "2", "--word_repr_size", "64",
                          "--entity_repr_size", "32", "--batch_size", "2048", "--window_size", "5", "--num_random_entities", "4",
                          "--seed", "3", "--update_method", "full_adam", "--nonlinearity", "hard_tanh", "--batch_normalization",
                          "--synthetic_num_words", "2000", "--synthetic_num_entities", "500", "--synthetic_num_batches", "20",
                          "--output", out], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    assert "Epoch #2" in res.stdout and "n-grams/second" in res.stdout
    W = np.load(out + "_2.word_representations-representations.npy")
    T = np.load(out + "_2.word_entity_mapping-transform.npy")
    assert W.shape == (2000, 64) and T.shape == (64, 32) and np.isfinite(W).all()


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [["--nonlinearity", "tanh"], ["--nonlinearity", "tanh", "--batch_normalization"],
                                   ["--nonlinearity", "tanh", "--bias_negative_samples"]], ids=["lse", "bn", "biased"])
def test_cli_check_gradients(extra):
    """cuNVSMTrainModel --check_gradients (cpp/main.cu:63,414-420; GradientCheckFn, cpp/gradient_check.cu:3-133): every
    parameter of a tiny model against central differences of the cost on every batch, negatives replayed from the saved
    RNG state; the CLI aborts when a gradient points the wrong way or is off by more than 10 %."""
    _build()
    res = subprocess.run([os.path.join(CPP, "cuNVSMTrainModel"), "--check_gradients", "--num_epochs", "1", "--word_repr_size", "8",
                          "--entity_repr_size", "6", "--batch_size", "1024", "--window_size", "3", "--num_random_entities", "2",
                          "--seed", "5", "--update_method", "sgd", "--learning_rate", "0.05", "--synthetic_num_words", "30",
                          "--synthetic_num_entities", "20", "--synthetic_num_batches", "2", "--gemm", "fp32"] + extra,
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("Gradient check:")]
    assert len(lines) == 2, res.stdout
    for l in lines:
        checked = int(l.split()[2])
        assert checked >= 200, l          # 30*8 + 20*6 + 8*6 + 6 = 414 parameters, most above the noise floor
    assert "Epoch #1" in res.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("flag", ["--entity_similarity_weight", "--term_similarity_weight"])
def test_cli_trains_mixture_objective(tmp_path, flag):
    """cuNVSMTrainModel with a mixture weight selects TextEntityEntityEntity / TextEntityTermTerm (cpp/main.cu:729-757)."""
    _build()
    res = subprocess.run([os.path.join(CPP, "cuNVSMTrainModel"), "--num_epochs", "2", "--word_repr_size", "64",
                          "--entity_repr_size", "32", "--batch_size", "1024", "--window_size", "5", "--num_random_entities", "4",
                          "--seed", "3", "--update_method", "full_adam", "--nonlinearity", "tanh", flag, "0.25",
                          "--synthetic_num_words", "2000", "--synthetic_num_entities", "500", "--synthetic_num_batches", "10"],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    assert "Epoch #2" in res.stdout
    import re
    costs = [float(x) for x in re.findall(r"mean cost ([0-9.eE+-]+)", res.stdout)]
    assert len(costs) == 2 and np.isfinite(costs).all() and costs[1] < costs[0]


@pytest.mark.gpu
def test_cli_ngram_file_source_async_prefetch(tmp_path):
    """NGramFileSource + AsyncSource (the reference's prefetch contract, cpp/data_async.cpp) behind the CLI: two epochs
    over a pre-tokenised n-gram file, shuffled with the shared minstd_rand0 engine. The host-sampler run and the
    device-sampler run consume that engine in the same order (init -> shuffle -> negatives), so their costs agree."""
    _build()
    rng = np.random.default_rng(0)
    n, N, V, D = 4, 8192 + 300, 500, 120          # 8 full batches of 1024 + a partial one the CLI must skip
    words = rng.integers(0, V, size=(N, n))
    docs = (words[:, 0] * 3 + words[:, 1]) % D     # learnable
    path = tmp_path / "ngrams.txt"
    with open(path, "w") as f:
        f.write("# entity w1 w2 w3 w4\n")
        for i in range(N):
            f.write("%d %s%s\n" % (docs[i], " ".join(map(str, words[i])), " | 1.5" if i % 7 == 0 else ""))
    outs = []
    for extra in (["--host_sampler"], []):
        res = subprocess.run([os.path.join(CPP, "cuNVSMTrainModel"), "--num_epochs", "2", "--word_repr_size", "32",
                              "--entity_repr_size", "16", "--batch_size", "1024", "--window_size", str(n),
                              "--num_random_entities", "3", "--seed", "11", "--update_method", "sgd", "--nonlinearity", "tanh",
                              "--gemm", "fp32", "--ngram_file", str(path), "--num_concurrent_batches", "3",
                              "--feature_weighting", "self_information", "--weighting", "inv_doc_frequency",
                              "--output", str(tmp_path / "model")] + extra,
                             capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr
        # the `_meta` file of the reference (cpp/main.cu:527-537): term / object mapping + corpus statistics
        meta = subprocess.run([os.path.join(CPP, "cuNVSMMeta"), "print", str(tmp_path / "model_meta")], check=True,
                              capture_output=True, text=True).stdout.strip().split("\n")
        assert meta[-1] == "total_terms %d" % (N * n)
        assert sum(l.startswith("term ") for l in meta) == words.max() + 1
        assert sum(l.startswith("object ") for l in meta) == docs.max() + 1
        assert meta[7] == "term 7 7 %d" % (words == 7).sum()
        os.remove(tmp_path / "model_meta")
        assert "|V|=%d |D|=%d" % (words.max() + 1, docs.max() + 1) in res.stdout
        assert res.stderr.count("Skipping Batch") == 2      # the partial batch of each epoch
        import re
        outs.append([float(x) for x in re.findall(r"mean cost ([0-9.eE+-]+)", res.stdout)])
    assert len(outs[0]) == 2 and np.isfinite(outs[0]).all()
    np.testing.assert_allclose(outs[0], outs[1], rtol=1e-6)   # identical batches, identical negatives


@pytest.mark.gpu
def test_cli_zipf_negatives_host_and_device_samplers_agree():
    """--negative_sampling_zipf installs InverseCdfLabelGenerator (the reference's LabelGenerator plug point,
    include/cuNVSM/labels.h:7-18): the host loop and the device sampler draw the same skewed negatives, so the
    two runs report the same costs; they differ from the uniform run."""
    _build()
    import re
    costs = {}
    for name, extra in (("host", ["--negative_sampling_zipf", "1.0", "--host_sampler"]),
                        ("device", ["--negative_sampling_zipf", "1.0"]), ("uniform", [])):
        res = subprocess.run([os.path.join(CPP, "cuNVSMTrainModel"), "--num_epochs", "2", "--word_repr_size", "32",
                              "--entity_repr_size", "16", "--batch_size", "1024", "--window_size", "4",
                              "--num_random_entities", "5", "--seed", "9", "--update_method", "adagrad", "--nonlinearity", "tanh",
                              "--gemm", "fp32", "--bias_negative_samples", "--synthetic_num_words", "800",
                              "--synthetic_num_entities", "3000", "--synthetic_num_batches", "6"] + extra,
                             capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr
        costs[name] = [float(x) for x in re.findall(r"mean cost ([0-9.eE+-]+)", res.stdout)]
        assert len(costs[name]) == 2 and np.isfinite(costs[name]).all()
    np.testing.assert_allclose(costs["host"], costs["device"], rtol=1e-6)
    assert abs(costs["host"][0] - costs["uniform"][0]) > 1e-4 * costs["uniform"][0]


@pytest.mark.gpu
def test_cli_mixture_with_similarity_file(tmp_path):
    """The TextEntityEntityEntity mixture fed like the reference feeds it (cpp/main.cu:279-305): n-gram source +
    similarity file resolved through the identifiers map, shuffled with the shared RNG and repeated for the length of
    the text epoch (RepeatingSource(-1)). Pairs naming an unknown entity are skipped."""
    _build()
    import re
    rng = np.random.default_rng(1)
    n, N, V, D = 4, 4096, 300, 90
    words = rng.integers(0, V, size=(N, n))
    docs = (words[:, 0] * 3 + words[:, 1]) % D
    words[0, 0], docs[0] = V - 1, D - 1
    ngrams = tmp_path / "ngrams.txt"
    with open(ngrams, "w") as f:
        for i in range(N):
            f.write("%d %s\n" % (docs[i], " ".join(map(str, words[i]))))
    pairs = tmp_path / "pairs.txt"
    with open(pairs, "w") as f:
        for i in range(1500):                      # fewer pairs than the text epoch needs: the source repeats
            f.write("%d %d %.2f\n" % (rng.integers(0, D), rng.integers(0, D), 0.5 + (i % 3)))
        f.write("%d 12 1.0\n" % (D + 5))           # unknown entity
    res = subprocess.run([os.path.join(CPP, "cuNVSMTrainModel"), "--num_epochs", "2", "--word_repr_size", "32",
                          "--entity_repr_size", "16", "--batch_size", "1024", "--window_size", str(n), "--num_random_entities", "3",
                          "--seed", "4", "--update_method", "full_adam", "--nonlinearity", "tanh", "--gemm", "fp32",
                          "--entity_similarity_weight", "0.3", "--ngram_file", str(ngrams), "--similarity_file", str(pairs),
                          "--host_sampler"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    assert "similarity file: 1500 pairs" in res.stdout and "not found; skipping pair" in res.stderr
    costs = [float(x) for x in re.findall(r"mean cost ([0-9.eE+-]+)", res.stdout)]
    assert len(costs) == 2 and np.isfinite(costs).all() and costs[1] < costs[0]
