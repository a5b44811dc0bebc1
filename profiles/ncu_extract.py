"""Reads an ncu CSV (`ncu --csv --metrics …` log, or `ncu -i rep --page raw --csv`) of the solve kernel captured by
profiles/profile_solve.py and writes profiles/traffic.json — the measured counters bench.py's roofline block uses.
usage: python profiles/ncu_extract.py <metrics.csv> <profile_solve_B.json> <tag>"""
import csv, json, os, sys

path, meta_path, tag = sys.argv[1], sys.argv[2], sys.argv[3]
meta = json.load(open(meta_path))
rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
hdr = rows[0]
vals = {}
if "Metric Name" in hdr:                                  # long format: one row per metric
    ni, vi, ki = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Kernel Name")
    for r in rows[1:]:
        if "agb_newton_solve" in r[ki]:
            vals[r[ni]] = float(r[vi].replace(",", ""))
else:                                                      # raw page: one column per metric, second row = units
    for r in rows[2:]:
        if any("agb_newton_solve" in c for c in r):
            for k, v in zip(hdr, r):
                try:
                    vals[k] = float(v.replace(",", ""))
                except ValueError:
                    pass
g = lambda k: vals.get(k)
steps = meta["newton_steps_per_launch"]
dfma, dadd, dmul = (g("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % o) or 0.0 for o in ("dfma", "dadd", "dmul"))
winst, tinst = g("smsp__inst_executed.sum"), g("smsp__thread_inst_executed.sum")
out = {
    "tag": tag, "config": meta["config"], "batch": meta["batch"], "newton_steps_of_captured_launch": steps,
    "dram_bytes_per_launch": (g("dram__bytes_read.sum") or 0.0) + (g("dram__bytes_write.sum") or 0.0),
    "dfma_thread_inst": dfma, "dadd_thread_inst": dadd, "dmul_thread_inst": dmul,
    "fp64_flops_per_newton_step": (2 * dfma + dadd + dmul) / steps if steps else None,
    "warp_inst_per_launch": winst, "warp_inst_per_newton_step": winst / steps if winst and steps else None,
    "lanes_per_inst": tinst / winst if winst and tinst else None,
    "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "fp64_pipe_busy_pct": g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active") or g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "registers_per_thread": g("launch__registers_per_thread"), "ctas_per_sm": g("launch__occupancy_limit_shared_mem") and min(
        x for x in (g("launch__occupancy_limit_shared_mem"), g("launch__occupancy_limit_registers"), g("launch__occupancy_limit_warps")) if x),
    "launch_ms_under_ncu": (g("gpu__time_duration.sum") or 0.0) / 1e6,
    "source": f"profiles/{tag}_ncu_metrics.csv (ncu --metrics, one launch of agb_newton_solve_kernel, config {meta['config']} batch {meta['batch']}); "
              "fp64 flops = 2*dfma + dadd + dmul thread instructions (pred on)",
}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
