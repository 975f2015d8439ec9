#!/usr/bin/env python
"""Condenses an .ncu-rep (read here, without a GPU) into the small text summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/k1.ncu-rep [--structures N] > profiles/r2_k1_ncu_summary.txt

`--structures N` records how many structures one profiled launch processed (bench.py scales the DRAM traffic by it).
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__block_size",
    "launch__grid_size",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def table(rep, peak_gbs):
    """One line per profiled launch: the digest kept as profiles/*_all_kernels_ncu_table.txt."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}

    def num(d, key, factor=None):
        v = float(d[col[key]].replace(",", ""))
        u = units[col[key]]
        return v * (factor[u] if factor else 1.0)

    print(f"# One `ncu --set full --clock-control none` capture of every kernel family at its BASELINE.json size")
    print(f"# (tools/all_kernels_probe.py; cold caches, serialised: compare shares and byte counts, not absolute times).")
    print(f"# GB/s = (dram read + write) / duration; frac = GB/s / {peak_gbs} (MEASURED_PEAKS.json).  Outputs of the small")
    print(f"# kernels stay in the 126 MB L2 (wr_MB ~ 0), so their DRAM figures understate the bytes they move.")
    print(f"{'kernel':52s} {'time_us':>8s} {'rd_MB':>9s} {'wr_MB':>9s} {'GB/s':>7s} {'frac':>6s} {'dram%':>6s} {'issue%':>6s} "
          f"{'warps%':>6s} {'fma_pipe%':>9s} {'regs':>5s}")
    for d in data:
        name = d[col["Kernel Name"]]
        name = name.replace("void ", "").replace("unnamed>::", "").replace("ps::<unnamed>::", "")
        name = name.split("(")[0] if "<" not in name else name[:name.rfind(">") + 1].split("(const")[0]
        t = num(d, "gpu__time_duration.sum", tscale)
        rd = num(d, "dram__bytes_read.sum", scale) / 1e6
        wr = num(d, "dram__bytes_write.sum", scale) / 1e6
        gbs = (rd + wr) * 1e6 / (t * 1e-6) / 1e9
        fma = d[col["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]] if \
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active" in col else "nan"
        print(f"{name[:52]:52s} {t:8.1f} {rd:9.1f} {wr:9.1f} {gbs:7.0f} {gbs / peak_gbs:6.2f} "
              f"{float(d[col['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']]):6.1f} "
              f"{float(d[col['smsp__issue_active.avg.pct_of_peak_sustained_active']]):6.1f} "
              f"{float(d[col['sm__warps_active.avg.pct_of_peak_sustained_active']]):6.1f} {float(fma):9.1f} "
              f"{d[col['launch__registers_per_thread']]:>5s}")


def main():
    rep = sys.argv[1]
    if "--table" in sys.argv:
        table(rep, float(sys.argv[sys.argv.index("--table") + 1]))
        return
    structures = int(sys.argv[sys.argv.index("--structures") + 1]) if "--structures" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary of {rep} ({len(data)} profiled launch(es)); ncu --set full --clock-control none")
    if structures is not None:
        print(f"# structures per launch: {structures}")
    for d in data:
        print(f"\nkernel: {d[col['Kernel Name']]}")
        for k in KEYS:
            if k in col:
                print(f"  {k:72s} {d[col[k]]:>18s} {units[col[k]]}")
        print("  warp stall reasons (average warps stalled per issue-active cycle):")
        stalls = []
        for h, i in col.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(d[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        for v, name in sorted(stalls, reverse=True)[:10]:
            print(f"    {name:28s} {v:8.3f}")
        t_ms = float(d[col["gpu__time_duration.sum"]]) * (1e-3 if units[col["gpu__time_duration.sum"]] == "us" else 1.0)
        wr = float(d[col["dram__bytes_write.sum"]])
        wu = units[col["dram__bytes_write.sum"]]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(wu, 1.0)
        print(f"  derived: DRAM write rate under ncu = {wr * scale / (t_ms / 1e3) / 1e9:.1f} GB/s "
              f"(ncu timings are cold-cache / serialised; bench.py carries the number of record)")


if __name__ == "__main__":
    main()
