"""ctypes binding of libnvsm_b200.so (C ABI: include/nvsm_b200.h).

There is no fallback: if the shared library is missing or cannot be loaded this module
raises at first use, and every compute entry point fails without a CUDA device.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NVSM_LIB_PATH") or os.path.join(HERE, "libnvsm_b200.so")   # (override: kernel-variant experiments)

# every symbol include/nvsm_b200.h declares
SYMBOLS = [
    "nvsm_last_error", "nvsm_version", "nvsm_host_alloc", "nvsm_host_free", "nvsm_create", "nvsm_destroy", "nvsm_set_stream", "nvsm_synchronize",
    "nvsm_initialize", "nvsm_tensor_size", "nvsm_get_tensor", "nvsm_set_tensor", "nvsm_generate_labels",
    "nvsm_compute_cost", "nvsm_wait_upload", "nvsm_compute_gradients", "nvsm_update", "nvsm_get_cost", "nvsm_read_cost", "nvsm_read_cost_f64",
    "nvsm_scaled_regularization_lambda", "nvsm_train_step", "nvsm_stage_batch", "nvsm_compute_cost_staged",
    "nvsm_train_step_staged", "nvsm_infer", "nvsm_increment_parameter", "nvsm_set_profiling", "nvsm_num_phases",
    "nvsm_phase_name", "nvsm_get_phase_ms", "nvsm_get_timeline", "nvsm_reset_phase_ms", "nvsm_kernel_launches", "nvsm_comm_unique_id",
    "nvsm_comm_init", "nvsm_comm_set_sparse_mode", "nvsm_comm_peer_export", "nvsm_comm_peer_import", "nvsm_comm_peer_status", "nvsm_comm_peer_disable", "nvsm_similarity_compute_cost", "nvsm_similarity_get_cost",
    "nvsm_similarity_scaled_regularization_lambda", "nvsm_test_gemm_tc", "nvsm_bench_gemm_tc", "nvsm_bench_memory", "nvsm_sampler_seed", "nvsm_sampler_state",
    "nvsm_ops_create", "nvsm_ops_destroy", "nvsm_ops_synchronize", "nvsm_ops_kernel_launches", "nvsm_dev_malloc", "nvsm_dev_free",
    "nvsm_dev_upload", "nvsm_dev_download", "nvsm_dev_copy", "nvsm_dev_fill", "nvsm_op_average_representations", "nvsm_op_project",
    "nvsm_op_activation", "nvsm_op_batchnorm_forward", "nvsm_op_batchnorm_backward", "nvsm_op_update_dense",
    "nvsm_op_representations_update", "nvsm_op_transform_update", "nvsm_updater_create", "nvsm_updater_destroy",
    "nvsm_updater_update_representations", "nvsm_updater_update_transform", "nvsm_updater_state",
    "nvsm_step_sampled", "nvsm_get_entity_ids", "nvsm_generate_labels_device", "nvsm_generate_labels_cdf", "nvsm_sampler_set_cdf",
]


class NvsmConfig(ctypes.Structure):
    _fields_ = [
        ("num_words", ctypes.c_long), ("num_entities", ctypes.c_long),
        ("word_repr_size", ctypes.c_int), ("entity_repr_size", ctypes.c_int),
        ("nonlinearity", ctypes.c_int), ("batch_normalization", ctypes.c_int),
        ("clip_sigmoid", ctypes.c_int), ("bias_negative_samples", ctypes.c_int),
        ("l2_normalize_phrase_reprs", ctypes.c_int), ("l2_normalize_entity_reprs", ctypes.c_int),
        ("update_method", ctypes.c_int), ("adam_mode", ctypes.c_int),
        ("num_random_entities", ctypes.c_int), ("max_batch_size", ctypes.c_int),
        ("window_size", ctypes.c_int), ("regularization_lambda", ctypes.c_float),
        ("device", ctypes.c_int), ("gemm_mode", ctypes.c_int), ("num_batch_slots", ctypes.c_int),
        ("objective", ctypes.c_int), ("text_entity_weight", ctypes.c_float), ("similarity_weight", ctypes.c_float),
        ("max_similarity_batch_size", ctypes.c_int), ("reserved", ctypes.c_int * 3),
    ]


_lib = None


def load():
    """Load libnvsm_b200.so and declare the prototypes. Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libnvsm_b200.so is not built (%s). Run `python -m cunvsm_b200.build`; there is no CPU "
            "or PyTorch fallback for the NVSM step." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, cl, ci, cf = ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_float
    pl, pf = ctypes.POINTER(ctypes.c_long), ctypes.POINTER(ctypes.c_float)
    pul = ctypes.POINTER(ctypes.c_ulong)
    cs = ctypes.c_char_p

    def f(name, argtypes, restype=ci):
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = restype

    f("nvsm_last_error", [], cs)
    f("nvsm_version", [])
    f("nvsm_host_alloc", [ctypes.POINTER(vp), ctypes.c_ulong])
    f("nvsm_host_free", [vp])
    f("nvsm_create", [ctypes.POINTER(NvsmConfig), ctypes.POINTER(vp)])
    f("nvsm_destroy", [vp], None)
    f("nvsm_set_stream", [vp, vp])
    f("nvsm_synchronize", [vp])
    f("nvsm_initialize", [vp, pul])
    f("nvsm_tensor_size", [vp, cs], cl)
    f("nvsm_get_tensor", [vp, cs, pf, cl])
    f("nvsm_set_tensor", [vp, cs, pf, cl])
    f("nvsm_generate_labels", [pl, cl, cl, cl, pul, pl])
    f("nvsm_compute_cost", [vp, pl, pf, pl, pf, cl])
    f("nvsm_wait_upload", [vp])
    f("nvsm_compute_gradients", [vp])
    f("nvsm_update", [vp, cf, cf])
    f("nvsm_get_cost", [vp, pf])
    f("nvsm_read_cost", [vp, ci, pf])
    f("nvsm_scaled_regularization_lambda", [vp], cf)
    f("nvsm_train_step", [vp, pl, pf, pl, pf, cl, cf])
    f("nvsm_stage_batch", [vp, ci, pl, pf, pl, pf, cl])
    f("nvsm_compute_cost_staged", [vp, ci])
    f("nvsm_train_step_staged", [vp, ci, cf])
    f("nvsm_infer", [vp, pl, cl, cl, pf])
    f("nvsm_increment_parameter", [vp, cs, cl, cf])
    f("nvsm_set_profiling", [vp, ci])
    f("nvsm_num_phases", [])
    f("nvsm_phase_name", [ci], cs)
    f("nvsm_get_phase_ms", [vp, pf, ci])
    f("nvsm_reset_phase_ms", [vp])
    f("nvsm_get_timeline", [vp, ctypes.POINTER(ci), pf, pf, ci])
    f("nvsm_kernel_launches", [vp], cl)
    f("nvsm_test_gemm_tc", [vp, ci, ci, ci, ci, pf, pf, pf, cf, pf, ci])
    f("nvsm_bench_gemm_tc", [vp, ci, ci, ci, ci, ci, ci, ci, pf])
    f("nvsm_read_cost_f64", [vp, ci, ctypes.POINTER(ctypes.c_double)])
    f("nvsm_bench_memory", [vp, ci, cl, ci, ci, cl, ci, pf])
    f("nvsm_sampler_seed", [vp, ctypes.c_ulong])
    f("nvsm_sampler_state", [vp, pul])
    f("nvsm_step_sampled", [vp, pl, pf, pl, pf, cl, cf, ci])
    f("nvsm_get_entity_ids", [vp, pl, cl])
    f("nvsm_generate_labels_device", [vp, pl, cl, cl, cl, pul, pl])
    pd = ctypes.POINTER(ctypes.c_double)
    f("nvsm_generate_labels_cdf", [pl, cl, cl, pd, cl, pul, pl])
    f("nvsm_sampler_set_cdf", [vp, pd, cl])
    f("nvsm_comm_unique_id", [ctypes.c_char_p])
    f("nvsm_comm_init", [vp, ctypes.c_char_p, ci, ci])
    f("nvsm_comm_set_sparse_mode", [vp, ci])
    f("nvsm_comm_peer_export", [vp, ctypes.c_char_p])
    f("nvsm_comm_peer_import", [vp, ctypes.c_char_p])
    f("nvsm_comm_peer_status", [vp, ctypes.POINTER(ci), ctypes.POINTER(ci)])
    f("nvsm_comm_peer_disable", [vp])
    f("nvsm_similarity_compute_cost", [vp, pl, pf, cl])
    f("nvsm_similarity_get_cost", [vp, pf])
    f("nvsm_similarity_scaled_regularization_lambda", [vp], cf)
    _lib = L
    return L


class NvsmError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise NvsmError(load().nvsm_last_error().decode("utf-8", "replace"))
