"""Per-kernel key metrics from an ncu --set full report (raw page) as a markdown table."""
import csv, subprocess, sys
KEYS = [("gpu__time_duration.sum", "dur_us"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_%"),
        ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor_inst_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_%"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "smem_dyn"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
        ("smsp__inst_executed.sum", "warp_insts")]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("### %s\n" % rep)
    print("| kernel | grid | block | " + " | ".join(n for _, n in KEYS) + " |")
    print("|---|---|---|" + "---|" * len(KEYS))
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        vals = []
        for k, _ in KEYS:
            if k in ix:
                v = r[ix[k]]
                u = units[ix[k]]
                try:
                    f = float(v.replace(",", ""))
                    if k == "gpu__time_duration.sum":
                        f = f / 1e3 if u == "ns" else (f * 1e3 if u == "ms" else f)
                    if u in ("Mbyte", "Kbyte", "Gbyte", "byte") :
                        f *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
                        v = "%.3g MB" % (f / 1e6)
                    else:
                        v = "%.4g" % f
                except ValueError:
                    pass
                vals.append(v)
            else:
                vals.append("-")
        print("| %s | %s | %s | " % (name, r[ix["Grid Size"]], r[ix["Block Size"]]) + " | ".join(vals) + " |")
    print()
