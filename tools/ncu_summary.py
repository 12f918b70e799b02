#!/usr/bin/env python
"""Condense an ncu capture (--set full) of one kernel into a small JSON + the raw-page CSV under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r02_evaluate_kernel [kernel-name-substring] [build-id]

With a kernel-name substring the first launch whose name contains it is condensed (a report may hold several kernels); the build
id (``python -c "from dlii_jssp_b200 import build; print(build.library_id())"`` of the library that was profiled) is stored in
the JSON: bench.py only attaches a capture to a library with the same build id.

bench.py reads the JSON (if present) to fill roofline.traffic and roofline.executed - measured under the profiler, so they
are evidence about the kernel, never a bench value."""
import csv, io, json, subprocess, sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__warps_eligible.avg.per_cycle_active": "eligible_warps_per_cycle",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_warp_inst",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "launch__waves_per_multiprocessor": "waves_per_sm",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_kb",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed": "thread_dadd_per_cycle",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed": "thread_dmul_per_cycle",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed": "thread_dfma_per_cycle",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed": "thread_fadd_per_cycle",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed": "thread_fmul_per_cycle",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed": "thread_ffma_per_cycle",
    "sm__cycles_elapsed.avg": "sm_cycles_elapsed",
    "sm__cycles_elapsed.avg.per_second": "sm_ghz",
}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "ms": 1e3, "s": 1e6, "second": 1e6}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    want = sys.argv[3] if len(sys.argv) > 3 else ""
    build_id = sys.argv[4] if len(sys.argv) > 4 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kcol = hdr.index("Kernel Name")
    vals = next((r for r in rows[2:] if want in r[kcol]), None)
    if vals is None:
        raise SystemExit(f"no kernel matching {want!r} in {rep}: " + ", ".join(sorted({r[kcol] for r in rows[2:]})))
    d = {"kernel": vals[kcol], "source": rep.split("/")[-1], "build_id": build_id}
    for i, h in enumerate(hdr):
        if h in KEYS:
            v = float(vals[i].replace(",", ""))
            d[KEYS[h]] = v * UNIT_SCALE.get(units[i], 1.0) if KEYS[h] in ("dram_read", "dram_write", "duration_us") else v
    d["dram_traffic_bytes"] = d.get("dram_read", 0.0) + d.get("dram_write", 0.0)
    cyc = d.get("sm_cycles_elapsed", 0.0)
    d["executed_fp64_flops"] = (d.get("thread_dadd_per_cycle", 0) + d.get("thread_dmul_per_cycle", 0) + 2 * d.get("thread_dfma_per_cycle", 0)) * cyc
    d["executed_fp32_flops"] = (d.get("thread_fadd_per_cycle", 0) + d.get("thread_fmul_per_cycle", 0) + 2 * d.get("thread_ffma_per_cycle", 0)) * cyc
    json.dump(d, open(out + "_ncu.json", "w"), indent=1)
    with open(out + "_ncu.csv", "w") as f:
        w = csv.writer(f)
        for i, h in enumerate(hdr):
            w.writerow([h, units[i], vals[i]])
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main()
