"""Extract the known-answer literals of the reference's own gtest suites into JSON.

Run once in the build container (needs /root/reference, which does not exist on
the GPU box); the output reference_goldens.json is committed next to this script.
Only numeric literals are extracted (file:line ranges recorded in the JSON);
no reference code is copied.

    python tests/golden/extract_reference_goldens.py
"""
import json
import os
import re

REF = "/root/reference"
NUM = r"FPHelper<FloatT>::eq\(\s*(-?[0-9][0-9.eE+-]*)\s*\)"


def literals(path, first, last):
    with open(os.path.join(REF, path)) as f:
        lines = f.readlines()[first - 1:last]
    return [float(x) for x in re.findall(NUM, "".join(lines))]


def main():
    out = {}
    mt = "cpp/model_tests.cu"
    out["transform_backward"] = {
        "source": mt + ":341-466",
        "config": {"seed": 10, "num_words": 5, "num_entities": 3, "word_repr_size": 2,
                   "entity_repr_size": 3, "bias_negative_samples": True, "num_random_entities": 10,
                   "regularization_lambda": 0.01, "batch_size": 32, "window_size": 2,
                   "feature_value": 2, "label": 1},
        "grad_transform": literals(mt, 377, 386),
        "grad_bias": literals(mt, 388, 395),
        "grad_phrase": literals(mt, 397, 465),
    }
    out["transform_bn_forward"] = {
        "source": mt + ":468-521",
        "epsilon": 1e-5,
        "output": literals(mt, 508, 521),
    }
    ct = "cpp/cudnn_utils_tests.cu"
    out["bn_forward_backward"] = {
        "source": ct + ":115-177",
        "epsilon": 1e-5,
        "input": [1.0, 2.0, 3.0, 5.0, 10.0, 20.0],
        "grad_output": [0.25, -0.1, 0.3, 1.0, 0.005, -0.5],
        "grad_bias": literals(ct, 163, 167),
        "grad_input": literals(ct, 169, 176),
    }
    ut = "cpp/updates_tests.cu"
    out["adam_transform"] = {
        "source": ut + ":299-425",
        "epsilon": 1e-5,
        "grad_bias_t1": literals(ut, 352, 354),
        "m_bias_t1": literals(ut, 358, 360),
        "v_bias_t1": literals(ut, 364, 366),
        "grad_bias_t2": literals(ut, 409, 411),
        "m_bias_t2": literals(ut, 415, 417),
        "v_bias_t2": literals(ut, 421, 423),
    }
    cu = "cpp/cuda_utils_tests.cu"
    with open(os.path.join(REF, cu)) as f:
        block = "".join(f.readlines()[81:92])          # the ElementsAre(...) of grad_input, :82-92
    out["normalizer"] = {
        "source": cu + ":51-92",
        "input": [[1, 2, 3, 4, 5], [6, 7, 8, 9, 10]],
        "grad_output": [[10000, 10001, 10002, 10003, 10004], [10005, 10006, 10007, 10008, 10009]],
        "grad_input": [float(x) for x in re.findall(r"(-?[0-9]+\.[0-9]+)", block)],
    }
    assert len(out["normalizer"]["grad_input"]) == 10, out["normalizer"]["grad_input"]
    assert len(out["transform_backward"]["grad_transform"]) == 6
    assert len(out["transform_backward"]["grad_bias"]) == 3
    assert len(out["transform_backward"]["grad_phrase"]) == 64
    assert len(out["transform_bn_forward"]["output"]) == 10
    assert len(out["bn_forward_backward"]["grad_bias"]) == 3
    assert len(out["bn_forward_backward"]["grad_input"]) == 6
    for k in ("grad_bias_t1", "m_bias_t1", "v_bias_t1", "grad_bias_t2", "m_bias_t2", "v_bias_t2"):
        assert len(out["adam_transform"][k]) == 3, (k, out["adam_transform"][k])
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "reference_goldens.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote reference_goldens.json")


if __name__ == "__main__":
    main()
