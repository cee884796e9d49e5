#!/usr/bin/env python
"""Per-layer table from an `ncu --page raw --csv` export of the layer-kernel launches of one evaluation step
(tools/gpu_evidence.sh): duration, tensor-pipe utilisation, executed tensor FLOPs, DRAM bytes, UTCMMA operand wavefronts.

    python tools/ncu_layer_table.py gpurun_out/evidence/prof_layers_raw.csv [--snippets 250] > profiles/rNN_ncu_layer_table.md
"""
import argparse
import csv
import sys

LAYERS = ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv4_1", "conv4_2", "conv4_3", "conv5_1",
          "conv5_2", "conv5_3", "FC1", "FC2", "FC3"]
# algorithmic GFLOP per snippet; conv1_1 differs per stream
GF = {"conv1_1": (0.1734, 1.1561), "conv1_2": 3.6994, "conv2_1": 1.8497, "conv2_2": 3.6994, "conv3_1": 1.8497, "conv3_2": 3.6994,
      "conv3_3": 3.6994, "conv4_1": 1.8497, "conv4_2": 3.6994, "conv4_3": 3.6994, "conv5_1": 0.9248, "conv5_2": 0.9248, "conv5_3": 0.9248,
      "FC1": 0.2055, "FC2": 0.0336, "FC3": 0.0021}


def num(v):
    try:
        return float(str(v).replace(",", ""))
    except ValueError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--snippets", type=int, default=250)
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    col = {n: i for i, n in enumerate(names)}

    def get(r, key):
        i = col.get(key)
        return num(r[i]) if i is not None and i < len(r) else None

    def unit(key):
        i = col.get(key)
        return units[i] if i is not None else ""

    def scaled(r, key, want):
        """value converted to `want` (ns/us/ms -> us; byte/Kbyte/Mbyte/Gbyte -> MB)"""
        v, u = get(r, key), unit(key).lower()
        if v is None:
            return None
        if want == "us":
            return v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "msecond": 1e3, "ms": 1e3, "nsecond": 1e-3, "second": 1e6, "s": 1e6}.get(u, 1.0)
        if want == "MB":
            return v * {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3, "tbyte": 1e6}.get(u, 1e-6)
        return v

    data = [r for r in data if len(r) > col["Kernel Name"]]
    # the capture window may start a few launches before the step's first layer: align on the first-layer kernel (the
    # only instances with CK = 16 / 32: conv_tc_kernel<64, 16, ..> spatial, <64, 32, ..> temporal, or the fused gather kernel)
    import re
    first = next((i for i, r in enumerate(data) if re.search(r"conv_tc_kernel<64, (16|32),|conv1_fused", r[col["Kernel Name"]])), 0)
    lead = data[:first]
    data = data[first:]
    for r in lead:
        print(f"(before the step's first layer: {r[col['Kernel Name']].split('(')[0]}, {scaled(r, 'gpu__time_duration.sum', 'us'):.1f} us -- tail of the previous stream)")
    print()
    print("| # | stream | layer | kernel | grid | duration us | algorithmic TFLOP/s | tensor pipe active % | executed TFLOP (tensor path) | DRAM read MB | "
          "DRAM write MB | UTCMMA A wavefronts | UTCMMA B wavefronts (1cta / 2cta) | SM clock MHz |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    tot = {}
    for k, r in enumerate(data):
        stream = "spatial" if k < 16 else "temporal"      # launch 16 (if present) is the temporal stream's conv1_1
        if k >= 17:
            break
        layer = LAYERS[k % 16]
        kname = r[col["Kernel Name"]].split("(")[0].replace("va::", "")
        dur = scaled(r, "gpu__time_duration.sum", "us")
        gf = GF[layer]
        gf = gf[0 if stream == "spatial" else 1] if isinstance(gf, tuple) else gf
        tfl = gf * a.snippets / dur / 1e3 if dur else None           # GFLOP / us = PFLOP/s -> x1e3 TFLOP/s ... (GF*n)/(us*1e-6)/1e12*1e9
        tfl = gf * a.snippets * 1e9 / (dur * 1e-6) / 1e12 if dur else None
        pipe = get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
        if pipe is None:
            pipe = get(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed")
        ops = get(r, "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum")
        rd, wr = scaled(r, "dram__bytes_read.sum", "MB"), scaled(r, "dram__bytes_write.sum", "MB")
        wa = get(r, "l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_a.sum")
        wb1 = get(r, "l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_b_scope_1cta.sum")
        wb2 = get(r, "l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_b_scope_2cta.sum")
        clk = get(r, "sm__cycles_elapsed.max")
        mhz = clk / dur if clk and dur else None
        grid = r[col["Grid Size"]] if "Grid Size" in col else ""
        f = lambda v, p=1: "" if v is None else (f"{v:.{p}f}")
        print(f"| {k} | {stream} | {layer} | {kname} | {grid} | {f(dur)} | {f(tfl, 0)} | {f(pipe)} | {f(ops / 1e12 if ops else None, 3)} | {f(rd)} | {f(wr)} | "
              f"{f(wa, 0)} | {f(wb1, 0)} / {f(wb2, 0)} | {f(mhz, 0)} |")
        t = tot.setdefault(stream, {"dur": 0.0, "pipe_w": 0.0, "rd": 0.0, "wr": 0.0})
        if dur:
            t["dur"] += dur
            t["pipe_w"] += (pipe or 0.0) * dur
        t["rd"] += rd or 0.0
        t["wr"] += wr or 0.0
    print()
    for stream, t in tot.items():
        gfs = sum((v[0 if stream == "spatial" else 1] if isinstance(v, tuple) else v) for v in GF.values())
        print(f"**{stream}**: {t['dur']:.0f} us for {a.snippets} snippets = {gfs * a.snippets * 1e9 / (t['dur'] * 1e-6) / 1e12:.0f} algorithmic TFLOP/s; "
              f"duration-weighted tensor-pipe activity {t['pipe_w'] / max(t['dur'], 1e-9):.1f} %; DRAM {t['rd']:.0f} MB read + {t['wr']:.0f} MB written.")


if __name__ == "__main__":
    main()
