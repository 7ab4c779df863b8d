#!/usr/bin/env python3
"""Turn the CSV exports of tools/ncu_capture_r02.sh (gpurun_out/<tag>_*) into the committed evidence under profiles/:
  <out>_<kernel>_raw.csv / _details.csv   copies of the ncu pages
  <out>_<kernel>_stalls.txt               stall-reason totals and hottest SASS windows (tools/ncu_source_summary.py)
  <out>_sass_evidence.txt                 counts of the Blackwell / Hopper mnemonics per kernel (UBLKCP = cp.async.bulk through the TMA unit, FFMA2 / FADD2 =
                                          packed fp32, SYNCS = mbarrier, UCGABAR = cluster barrier, ...) from the SASS pages
  ncu_traffic.json                        dram__bytes_read.sum + dram__bytes_write.sum per launch and other headline counters (read by bench.py)
usage: tools/profiles_from_capture.py <tag> <out-prefix>"""
import collections
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag, out = sys.argv[1], sys.argv[2]
G, P = ROOT / "gpurun_out", ROOT / "profiles"
KERNELS = {"k_pcg": "k_pcg", "k_schur": "k_schur", "k_kkt": "k_kkt", "k_merit_ls8": "k_merit_ls<8>", "k_merit_ls1": "k_merit_ls<1>", "k_pcg_cluster": "k_pcg_cluster",
           "rt_k_kkt": "k_kkt<RtPlant<7>> (table-driven)", "rt_k_merit_ls8": "k_merit_ls<RtPlant<7>, 8> (table-driven)"}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__icc_request_hit_rate.pct", "sass__inst_executed_register_spilling", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
MNEMONICS = ["UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "UCGABAR_ARV", "UCGABAR_WAIT", "MAPA", "LDGSTS", "LDS", "STS", "SHFL", "BAR", "MEMBAR", "CCTL", "ST", "ACQBULK"]
traffic, evidence = {}, []
for short, name in KERNELS.items():
    raw = G / f"{tag}_{short}_raw.csv"
    if not raw.exists():
        continue
    for page in ("raw", "details"):
        shutil.copy(G / f"{tag}_{short}_{page}.csv", P / f"{out}_{short}_{page}.csv")
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {}
    for k in WANT:
        for i, h in enumerate(hdr):
            if h == k:
                v = vals[i].replace(",", "")
                try:
                    v = float(v)
                except ValueError:
                    pass
                m[k] = {"value": v, "unit": units[i]}
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
    dram = sum(m[k]["value"] * scale.get(m[k]["unit"], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in m)
    if short.startswith("rt_"):  # evidence of the run-time model kernels only: bench.py's traffic table is about the compiled kernels
        pass
    traffic[name] = {"dram_bytes_per_launch": dram, "counters": m, "capture": f"ncu --set full --clock-control none, one launch inside the bench workload (B=512)" if short != "k_pcg_cluster" else "ncu --set full, one launch inside BASELINE config 4 (iiwa14, N=128, B=1024)"}
    sass = G / f"{tag}_{short}_source_sass.csv"
    txt = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_source_summary.py"), str(sass), "14"], capture_output=True, text=True).stdout
    (P / f"{out}_{short}_stalls.txt").write_text(txt)
    with open(sass) as f:
        r = csv.reader(f)
        next(r)
        h2 = next(r)
        si, ei = h2.index("Source"), h2.index("Instructions Executed")
        static, dyn = collections.Counter(), collections.Counter()
        for row in r:
            if len(row) <= ei:
                continue
            toks = row[si].split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            base = op.split(".")[0].rstrip(";")
            static[base] += 1
            dyn[base] += int(row[ei] or 0)
    evidence.append(f"{name}: " + ", ".join(f"{mn} {static[mn]} static / {dyn[mn]} executed" for mn in MNEMONICS if static[mn]))
(P / "ncu_traffic.json").write_text(json.dumps(traffic, indent=1))
(P / f"{out}_sass_evidence.txt").write_text("SASS mnemonic counts per kernel (static instructions in the kernel / warp-level instructions executed in the captured launch), from the ncu source pages.\n"
                                            "UBLKCP = cp.async.bulk (bulk copy through the TMA unit), SYNCS = mbarrier arrive / try_wait, FFMA2 / FADD2 = packed fp32x2 (sm_100+), UCGABAR_* = cluster barrier,\n"
                                            "MAPA = distributed-shared-memory address mapping, LDGSTS = cp.async.\n\n" + "\n".join(evidence) + "\n")
for f in (f"{tag}_launches_bench.csv", f"{tag}_launches_bench_summary.txt"):
    if (G / f).exists():
        shutil.copy(G / f, P / f.replace(tag, out))
print("\n".join(evidence))
print({k: round(v["dram_bytes_per_launch"] / 1e6, 2) for k, v in traffic.items()})
