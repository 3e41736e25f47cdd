"""cunvsm_b200 — B200-native (sm_100a) NVSM/LSE training step behind cuNVSM's interface.

The package holds only what the hot path needs: ``csrc/`` (hand-written CUDA kernels and the
C ABI of include/nvsm_b200.h, built into ``libnvsm_b200.so``) and the host-side mirror of the
reference's Model / Batch / config surface (``model.py``).
"""
from .model import (ENTITY_ENTITY, TERM_TERM, TEXT_ENTITY, TEXT_ENTITY_ENTITY_ENTITY, TEXT_ENTITY_TERM_TERM,
                    MultiForwardResult, SimilarityBatch, SimilarityForwardResult, ADAGRAD, ADAM, BIAS, DENSE_UPDATE, DENSE_UPDATE_DENSE_VARIANCE, ENTITY_REPRS, GEMM_3XTF32,
                    GEMM_FP32, GEMM_TF32, HARD_TANH, NONLINEARITIES, RNG, SGD, SPARSE, SPARSE_ALLGATHER, SPARSE_LOCAL, TANH, TRANSFORM,
                    UPDATE_METHODS, WORD_REPRS, Batch, ForwardResult, Gradients, Model, ModelDesc, NvsmError,
                    TrainConfig, comm_unique_id, zipf_cdf)

__all__ = [n for n in dir() if not n.startswith("_")]
