"""Summarise ncu artefacts brought back in gpurun_out/ into small tracked files under profiles/.

    python profiles/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
    python profiles/summarize_ncu.py full gpurun_out/prof_r1_step.ncu-rep profiles/r1_kernels_full.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = [
    ("gpu__time_duration.sum", "time_us", 1e-3),
    ("dram__bytes_read.sum", "dram_rd_MB", None),
    ("dram__bytes_write.sum", "dram_wr_MB", None),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct", 1),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct", 1),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__grid_size", "grid", 1),
    ("launch__block_size", "block", 1),
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("nvsm::", "")
    return name if len(name) <= 120 else name[:117] + "..."


def to_mb(value, unit):
    v = float(value)
    u = unit.lower()
    scale = {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, None)
    return v * scale if scale is not None else v


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(short(r[k]), []).append(float(r[v].replace(",", "")) / 1e3)
    total = sum(sum(x) for x in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n\n")
        f.write("source: %s, %d launches\n\n| kernel | launches | avg us | total us | share |\n|---|---|---|---|---|\n" % (src, len(rows)))
        for name, xs in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("| %s | %d | %.1f | %.1f | %.1f%% |\n" % (name, len(xs), sum(xs) / len(xs), sum(xs), 100 * sum(xs) / total))


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, rows = rows[0], rows[1], rows[2:]
    k = hdr.index("Kernel Name")
    cols = []
    for metric, label, scale in KEEP:
        idx = [i for i, h in enumerate(hdr) if h == metric]
        if idx:
            cols.append((idx[0], label, scale))
    with open(dst, "w") as f:
        f.write("# ncu --set full, one capture per launch of a step (%s)\n\n" % src)
        f.write("| kernel | " + " | ".join(c[1] for c in cols) + " |\n|---|" + "---|" * len(cols) + "\n")
        for r in rows:
            vals = []
            for i, label, scale in cols:
                raw = r[i].replace(",", "")
                try:
                    if scale is None:
                        vals.append("%.1f" % to_mb(raw, units[i]))
                    elif label == "time_us":
                        u = units[i].lower()
                        t = float(raw) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                        vals.append("%.1f" % t)
                    else:
                        vals.append("%.1f" % float(raw) if "." in raw else raw)
                except ValueError:
                    vals.append(raw)
            f.write("| %s | %s |\n" % (short(r[k]), " | ".join(vals)))


PHASE_OF = [("gather_mean", "gather_mean"), ("score_", "score_loss_bwd"), ("adam_full_pull_kernel<4, 2, 1>", "update_entities"),
            ("adam_full_pull_kernel<4, 3, 0>", "update_words"), ("entity_scatter", "update_entities"),
            ("word_scatter", "update_words"), ("bn_backward_kernel", "bn_backward"), ("gemm_tc_kernel<1, 1", "gemm_grad_transform")]


def traffic(src, dst, workload="C2"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes) of the dominant kernel of each
    bench phase -> profiles/traffic.json (read by bench.py for roofline.traffic)."""
    import json
    import os
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, rows = rows[0], rows[1], rows[2:]
    k, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    acc = {}
    for r in rows:
        name = short(r[k])
        for pat, phase in PHASE_OF:
            if pat in name:
                b = (to_mb(r[ir].replace(",", ""), units[ir]) + to_mb(r[iw].replace(",", ""), units[iw])) * 1e6
                acc.setdefault(phase, []).append(b)
                break
    data = json.load(open(dst)) if os.path.exists(dst) else {}
    data[workload] = {ph: int(sum(v) / len(v)) for ph, v in acc.items()}
    data["_source"] = "ncu --set full --clock-control none capture %s (per launch, bytes)" % os.path.basename(src)
    json.dump(data, open(dst, "w"), indent=1)
    print(data)


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
