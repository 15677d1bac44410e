"""Development tool: turn gpurun_out/<tag>_launches.csv and <tag>_*.ncu-rep into the text summaries under profiles/.
usage: python tools/summarize_profiles.py r1b"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second"]


def launches():
    src = os.path.join(G, tag + "_launches.csv")
    if not os.path.exists(src):
        return
    lines = [l for l in open(src) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot, n = collections.OrderedDict(), 0.0, 0
    for row in r:
        if len(row) <= iv:
            continue
        v = float(row[iv].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(row[iu], 1)
        k = row[ik].split("(")[0]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
    out = ["# %s: ncu launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` (B200, B=16, 800x1344)" % tag,
           "# ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c %d --csv" % n,
           "# %d consecutive launches; cold-cache, serialised -> compare SHARES, not absolutes" % n,
           "# total device time of these launches: %.3f ms" % (tot / 1e6), "",
           "%-60s %6s %10s %7s %9s" % ("kernel", "n", "total_ms", "share", "avg_us")]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%-60s %6d %10.3f %6.1f%% %9.1f" % (k[:60], c, t / 1e6, 100 * t / tot, t / c / 1e3))
    open(os.path.join(P, tag + "_launches_summary.txt"), "w").write("\n".join(out) + "\n")


def full(rep, title):
    path = os.path.join(G, rep + ".ncu-rep")
    csv_path = os.path.join(G, rep + "_raw.csv")   # exported on the GPU box by tools/capture_profiles.sh
    if os.path.exists(csv_path) and os.path.getsize(csv_path) > 0:
        raw = open(csv_path).read()
    elif os.path.exists(path):
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        return None
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = ["# " + title, ""]
    for n, row in enumerate(rows[2:]):
        out.append("launch %d: %s" % (n, row[hdr.index("Kernel Name")]))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.append("  %-82s %s %s" % (k, row[i], units[i]))
        out.append("")
    open(os.path.join(P, rep + "_ncu.txt"), "w").write("\n".join(out))
    return hdr, units, rows[2:]


launches()
if tag >= "r1f":   # fp16 convolution family (forward, scaled-gradient dgrad, MN-major wgrad)
    r = full(tag + "_conv_f16", "%s: ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_kernel "
             "-s 8 -c 6 (bench.py --steps 2 --warmup 1): fp16-operand launches of the first backward -- adapter dgrads "
             "(fp16 ReLU mask + channel sums + scaled fp16 output), teacher dgrads; B=16, P=22400 px/img, 422.8 GFLOP "
             "each" % tag)
    full(tag + "_conv_fwd16", "%s: ncu --set full ... -k regex:conv3x3_tc_kernel -s 0 -c 4: the first four forward launches "
         "(fp16 operands; conv+ReLU outputs that feed convolutions are written as fp16 only)" % tag)
    full(tag + "_wgrad_f16", "%s: ncu --set full ... -k regex:conv3x3_wgrad_kernel -s 2 -c 2: wgrad on fp16 operands "
         "(MN-major, SWIZZLE_128B, 8x8 pixel chunks, CTA pairs), 422.8 GFLOP each" % tag)
    full(tag + "_hbm", "%s: ncu --set full ... HBM-bound kernels of the step" % tag)
    full(tag + "_tokenprog", "%s: ncu --set full ... -k regex:token_program -c 2: the first two persistent token programs of a "
         "step (label encoder forward incl. the K / V projections; relation projections)" % tag)
    if r:
        hdr, units, rows = r
        mul = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
        tot = []
        for row in rows:
            rd = float(row[hdr.index("dram__bytes_read.sum")]) * mul[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(row[hdr.index("dram__bytes_write.sum")]) * mul[units[hdr.index("dram__bytes_write.sum")]]
            tot.append(rd + wr)
        json.dump({"dram_bytes_per_launch": sum(tot) / len(tot), "per_launch": tot,
                   "source": "profiles/%s_conv_f16_ncu.txt (mean over the captured fp16 dgrad launches), ncu --set full, "
                             "B=16 800x1344" % tag,
                   "algorithmic_bytes_per_launch": "fp16 in (0.5 F1 = 183.5 MB) + fp32 out (367 MB) or fp16 out (183.5 MB)"},
                  open(os.path.join(P, tag + "_conv3x3_traffic_raw.json"), "w"), indent=1)
    print("ok")
    sys.exit(0)
r = full(tag + "_conv_fwd", "%s: ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_kernel -s 8 -c 3 "
         "(bench.py --steps 2 --warmup 1): the three adapter dgrad launches of the first backward (two read a ReLU mask "
         "and emit channel sums, one does not); B=16, P=22400 px/img, 422.8 GFLOP each" % tag)
full(tag + "_conv_f16", "%s: ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_kernel -s 0 -c 4 "
     "(bench.py --steps 2 --warmup 1): the first four forward launches, fp16 operands (kind::f16, 64 channels per "
     "k-block): student_proj_2D, local_inst_proj_2D (+fp16 copy of the output), refinement 0, refinement 3; "
     "422.8 GFLOP each" % tag)
full(tag + "_conv_wgrad", "%s: ncu --set full ... -k regex:conv3x3_wgrad_kernel -s 2 -c 2: wgrad launches (MN-major tf32, "
     "SWIZZLE_128B_BASE32B, CTA pairs), 422.8 GFLOP each" % tag)
full(tag + "_hbm", "%s: ncu --set full ... HBM-bound kernels of the step (GroupNorm / InstanceNorm-MSE / pooling / rendering)" % tag)
if r:
    hdr, units, rows = r
    row = rows[-1]
    mul = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
    rd = float(row[hdr.index("dram__bytes_read.sum")]) * mul[units[hdr.index("dram__bytes_read.sum")]]
    wr = float(row[hdr.index("dram__bytes_write.sum")]) * mul[units[hdr.index("dram__bytes_write.sum")]]
    json.dump({"dram_bytes_per_launch": rd + wr, "read": rd, "write": wr,
               "source": "profiles/%s_conv_fwd_ncu.txt last launch (dgrad without ReLU-mask read), ncu --set full, B=16 800x1344" % tag,
               "algorithmic_bytes_per_launch": 2 * 1024 * 22400 * 16},
              open(os.path.join(P, "conv3x3_traffic.json"), "w"), indent=1)
print("ok")
