#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the per-kernel table committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_f64   ->  profiles/r1_f64_kernels.md + .json
"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_wavefront_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_excess_wavefronts"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__block_size", "block"),
    ("launch__grid_size", "grid"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem_blocks"),
    ("launch__occupancy_limit_registers", "occ_lim_reg_blocks"),
    ("sm__cycles_elapsed.avg", "cycles"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    kernels = []
    for r in rows[2:]:
        k = {"kernel": r[idx["Kernel Name"]].split("(")[0].replace("void ", "")}
        for m, short in METRICS:
            if m in idx:
                try:
                    k[short] = float(r[idx[m]].replace(",", ""))
                except ValueError:
                    k[short] = r[idx[m]]
                k[short + "_unit"] = units[idx[m]]
        kernels.append(k)
    json.dump(kernels, open(out + "_kernels.json", "w"), indent=1)
    with open(out + "_kernels.md", "w") as f:
        f.write(f"ncu --set full --clock-control none, source: {rep} (one launch per kernel)\n\n")
        cols = [s for _, s in METRICS]
        f.write("| kernel | " + " | ".join(cols) + " |\n|---|" + "---|" * len(cols) + "\n")
        for k in kernels:
            f.write("| " + k["kernel"] + " | " + " | ".join(
                (f"{k[c]:.4g} {k.get(c + '_unit', '')}".strip() if isinstance(k.get(c), float) else str(k.get(c, ""))) for c in cols) + " |\n")
    print(open(out + "_kernels.md").read())


if __name__ == "__main__":
    main()
