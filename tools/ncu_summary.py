#!/usr/bin/env python
"""Turn `ncu --set full` reports (gpurun_out/*.ncu-rep) into the small, tracked summaries under profiles/:
one JSON per report (duration, DRAM bytes, throughputs, occupancy, top stall reasons, hottest SASS lines)
and profiles/ncu_traffic.json = measured DRAM bytes per launch keyed by the kernel names bench.py prints.

    python tools/ncu_summary.py [--outdir=DIR] gpurun_out/r1_*.ncu-rep
(the reports are tens of MB each: run this on the GPU box and bring back only the summaries)
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second", "sm__cycles_active.avg",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]

# report stem -> (bench.py kernel name, launches of that kernel that one bench-level call makes)
BENCH_NAMES = {
    "r1_gemm_fwd1_bits": ("gemm_bf16_tc M=65536 N=2560 K=2560 a_mn=0 b_mn=0 epi=1", 1),
    "r1_gemm_dZ1_bits_db1": ("gemm_bf16_tc M=65536 N=2560 K=768 a_mn=0 b_mn=1 epi=9", 1),
    "r1_gemm_dX_dDP_fused": ("gemm_bf16_tc M=65536 N=2560 K=2560 a_mn=0 b_mn=1 epi=8", 1),
    "r1_gemm_fwd2": ("gemm_bf16_tc M=65536 N=768 K=2560 a_mn=0 b_mn=0 epi=7", 1),
    "r1_gemm_dW1": ("gemm_bf16_tc M=2560 N=2560 K=65536 a_mn=1 b_mn=1 epi=4", 1),
    "r1_perturb_fwd_shared6_bf16": ("perturb_gate_fwd B=65536 D=2560 models=6 out=bf16", 1),
    "r1_gemm_dW2": ("gemm_bf16_tc M=768 N=2560 K=65536 a_mn=1 b_mn=1 epi=4", 1),
    "r1_cls_ce_pass2": ("cls_ce B=65536 H=768 models=6 bwd=2", 6),
    "r1_cls_ce_pass1": ("cls_ce B=65536 H=768 models=6 bwd=1", 6),
    "r2_gemm_x3_fwd1": ("gemm_bf16x3_tc M=65536 N=2560 K=2560 a_mn=0 b_mn=0 epi=4", 1),
    "r2_gemm_x3_dW1": ("gemm_bf16x3_tc M=2560 N=2560 K=65536 a_mn=1 b_mn=1 epi=4", 1),
    "r2_split3": ("split3 R=65536 C=2560 act=1 planes=1 out=0 mask=0", 1),
    "r2_wide_adam_W1": ("linear_adam B=8 N=2304 K=2304 models=128", 1),
    "r2_wide_dx_W1": ("linear_bwd_dx B=8 N=2304 K=2304 models=128", 1),
}


def _num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return x


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def summarise(rep):
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    out = []
    for r in raw[2:]:
        d = dict(zip(hdr, r))
        k = {"kernel": d.get("Kernel Name"), "metrics": {}}
        for m in METRICS:
            if m in d and d[m] not in ("", "n/a"):
                k["metrics"][m] = [_num(d[m]), units[hdr.index(m)]]
        stalls = {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: _num(d[h]) for h in hdr
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and d[h] not in ("", "n/a")}
        k["stall_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        out.append(k)
    src = page(rep, "source")
    if len(src) > 2:
        h = src[1]
        try:
            i_src, i_s = h.index("Source"), h.index("# Samples")
            rows = [r for r in src[2:] if len(r) > i_s and r[i_s].strip().isdigit()]
            tot = sum(int(r[i_s]) for r in rows) or 1
            top = sorted(rows, key=lambda r: -int(r[i_s]))[:8]
            out[0]["hottest_sass"] = [{"sass": " ".join(r[i_src].split())[:90], "samples_pct": round(100.0 * int(r[i_s]) / tot, 1)} for r in top]
        except ValueError:
            pass
    return out


def main():
    reps = [a for a in sys.argv[1:] if not a.startswith("--outdir=")]
    outdir = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--outdir=")), os.path.join(ROOT, "profiles"))
    os.makedirs(outdir, exist_ok=True)
    traffic_path = os.path.join(outdir, "ncu_traffic.json")
    try:
        traffic = json.load(open(traffic_path))
    except Exception:
        traffic = {}
    for rep in reps:
        stem = os.path.splitext(os.path.basename(rep))[0]
        s = summarise(rep)
        json.dump({"report": os.path.basename(rep), "command": "ncu --set full --clock-control none --import-source on (tools/kbench.py --only <kernel> --iters 1)",
                   "kernels": s}, open(os.path.join(outdir, stem + "_ncu.json"), "w"), indent=1)
        m = s[0]["metrics"]
        rd, wr = m.get("dram__bytes_read.sum"), m.get("dram__bytes_write.sum")
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        if rd and wr and stem in BENCH_NAMES:
            name, mult = BENCH_NAMES[stem]
            traffic[name] = round((rd[0] * scale[rd[1]] + wr[0] * scale[wr[1]]) * mult)
        print(stem, s[0]["kernel"][:60], m.get("gpu__time_duration.sum"), "dram rd/wr", rd, wr)
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
