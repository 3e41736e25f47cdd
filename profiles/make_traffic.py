"""profiles/traffic.json from `ncu --page raw --csv` dumps of one profiled step per workload
(scripts/profile_step.py under `ncu --set full --clock-control none --profile-from-start off`):

    python profiles/make_traffic.py C2=gpurun_out/prof_r2a_C2.csv C3=... C5=...

Per workload and step phase: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and SM <-> L2 bytes
(lts__t_sectors_srcunit_tex.sum x 32) of the kernel(s) of that phase, per launch. bench.py reads it for
roofline.traffic / roofline.l2."""
import csv
import json
import sys

PHASES = (("gather_mean", ("gather_mean",)), ("gemm_fwd", ("gemm_tc2_kernel", "gemm_tc_kernel<0, 0")),
          ("bn_stats", ("col_stats",)), ("score_loss_bwd", ("score_ring_kernel", "score_kernel")),
          ("bn_backward", ("bn_backward",)), ("gemm_grad_transform", ("gemm_tc_kernel<1, 1",)),
          ("gemm_grad_phrase", ("gemm_tc_kernel<0, 0",)),
          ("update_entities", ("adam_full_pull_kernel<4, 2, 1>", "adam_full_pull_kernel<4, 1, 1>", "sgd_pull_kernel<4, 2, 1>",
                               "sgd_pull_kernel<4, 1, 1>", "sgd_pull_sparse_kernel<4, 2, 1>", "sgd_pull_sparse_kernel<4, 1, 1>",
                               "row_meansq_act_kernel")),
          ("update_words", ("adam_full_pull_kernel<4, 3, 0>", "adam_full_pull_kernel<4, 1, 0>", "sgd_pull_kernel<4, 3, 0>",
                            "sgd_pull_kernel<4, 1, 0>", "sgd_pull_sparse_kernel<4, 3, 0>", "sgd_pull_sparse_kernel<4, 1, 0>",
                            "row_meansq_kernel", "word_scalar_scatter", "word_adagrad_coef")))


def to_bytes(value, unit):
    scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "sector": 32.0}[unit.lower()]
    return float(value.replace(",", "")) * scale


def main():
    out = {}
    for arg in sys.argv[1:]:
        name, path = arg.split("=")
        rows = list(csv.reader(open(path)))
        hdr, units, rows = rows[0], rows[1], rows[2:]
        k = hdr.index("Kernel Name")
        rd, wr, tex = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("lts__t_sectors_srcunit_tex.sum")
        dram, l2 = {}, {}
        seen_fwd = False
        for r in rows:
            kn = r[k]
            for phase, pats in PHASES:
                if not any(p in kn for p in pats):
                    continue
                if phase == "gemm_fwd" and "gemm_tc_kernel<0, 0" in kn:
                    if seen_fwd:
                        continue          # the second K-major GEMM of the step is grad_phrase
                    seen_fwd = True
                elif phase == "gemm_fwd":
                    seen_fwd = True
                elif phase == "gemm_grad_phrase" and not dram.get("gemm_fwd"):
                    continue
                dram[phase] = dram.get(phase, 0.0) + to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr])
                l2[phase] = l2.get(phase, 0.0) + to_bytes(r[tex], units[tex])
                break
        out[name] = {p: int(v) for p, v in dram.items()}
        out[name + "_l2"] = {p: int(v) for p, v in l2.items()}
    out["_source"] = ("ncu --set full --clock-control none --profile-from-start off on scripts/profile_step.py (one steady-state "
                      "step per workload): " + ", ".join(sys.argv[1:]) + "; X = dram__bytes_read.sum + dram__bytes_write.sum per "
                      "phase, X_l2 = lts__t_sectors_srcunit_tex.sum x 32 B (SM <-> L2), bytes per launch")
    json.dump(out, open("profiles/traffic.json", "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
