"""Condense an .ncu-rep into the JSON summary committed under profiles/ (read here, in the build
container, with the same ncu that captured it):

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r1p_x_ncu_summary.json --symbols 1073741824 \
        --capture "ncu --set full ... python bench.py --steps 2 --warmup 3" --workload "bench.py: ..."

`--symbols` = symbols coded per launch; adds the derived per-symbol figures.
"""
import argparse
import csv
import io
import json
import subprocess

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "sm__cycles_elapsed.avg", "smsp__warps_eligible.avg.per_cycle_active",
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--symbols", type=float, default=0)
    ap.add_argument("--capture", default="")
    ap.add_argument("--workload", default="")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    unit_of = dict(zip(hdr, units))
    kernels = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        k = {"Kernel Name": d.get("Kernel Name", "")}
        for m in METRICS:
            if m in d and d[m] != "":
                k[m] = d[m]
        stalls = []
        for h in hdr:
            if h.startswith(STALL_PREFIX) and h.endswith("_per_issue_active.ratio") and d.get(h, "") not in ("", "n/a"):
                stalls.append((float(d[h]), h[len(STALL_PREFIX):-len("_per_issue_active.ratio")]))
        k["top_stalls_per_issue"] = {name: round(v, 3) for v, name in sorted(stalls, reverse=True)[:6]}
        if args.symbols:
            inst = float(d.get("smsp__inst_executed.sum", 0) or 0)
            wf = float(d.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 0) or 0)
            k["derived_warp_instructions_per_symbol"] = round(inst / (args.symbols / 32), 2)
            k["derived_smem_wavefronts_per_warp_symbol"] = round(wf / (args.symbols / 32), 2)
        try:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
            rd = float(d["dram__bytes_read.sum"]) * scale.get(unit_of["dram__bytes_read.sum"], 1)
            wr = float(d["dram__bytes_write.sum"]) * scale.get(unit_of["dram__bytes_write.sum"], 1)
            k["derived_dram_bytes_per_launch"] = int(rd + wr)
        except Exception:  # noqa: BLE001
            pass
        kernels.append(k)
    out = {"capture": args.capture, "workload": args.workload, "units": {m: unit_of.get(m, "") for m in ["Kernel Name"] + METRICS if m in unit_of},
           "kernels": kernels}
    json.dump(out, open(args.out, "w"), indent=1)
    for k in kernels:
        print(k["Kernel Name"][:70], k.get("gpu__time_duration.sum"), k.get("derived_warp_instructions_per_symbol"), k.get("derived_dram_bytes_per_launch"))


if __name__ == "__main__":
    main()
