"""Summarise .ncu-rep files (ncu --set full) into one JSON: per launch the headline metrics the roofline numbers cite.
Usage: python tools/ncu_rep_summary.py out.json a.ncu-rep b.ncu-rep ..."""
import csv, io, json, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum",
        "smsp__inst_executed.sum", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_op_umma_cycles_active.avg.pct_of_peak_sustained_active"]
out = []
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[0]
    units = dict(zip(hdr, rows[1]))
    scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        for k in list(d):  # normalise to ns / bytes
            u = units.get(k, "")
            if u in scale and scale[u] != 1.0:
                try:
                    d[k] = repr(float(d[k].replace(",", "")) * scale[u])
                except ValueError:
                    pass
        rec = {"file": rep.split("/")[-1], "kernel": d.get("Kernel Name", "")[:90]}
        for k in KEYS:
            if k in d:
                try:
                    rec[k] = float(d[k].replace(",", ""))
                except ValueError:
                    rec[k] = d[k]
        for k, v in d.items():  # tensor-pipe metrics are named differently across chips: keep whatever mentions the tensor pipe
            if "pipe_tensor" in k and k not in rec and v not in ("", "n/a"):
                try:
                    rec[k] = float(v.replace(",", ""))
                except ValueError:
                    pass
        out.append(rec)
open(sys.argv[1], "w").write(json.dumps(out, indent=1) + "\n")
for r in out:
    t = r.get("gpu__time_duration.sum", 0)
    print(r["kernel"][:60], r.get("launch__grid_size"), "us", round(t / 1e3, 1), "dram MB", round((r.get("dram__bytes_read.sum", 0) + r.get("dram__bytes_write.sum", 0)) / 1e6, 1),
          "dram%", r.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), "tensor%", r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
          "issue%", r.get("smsp__issue_active.avg.pct_of_peak_sustained_active"))
