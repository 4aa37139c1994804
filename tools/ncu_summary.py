"""Summarise an .ncu-rep into profiles/: per-kernel CSV (time, DRAM bytes, DRAM %, tensor %, regs) and the
per-launch DRAM traffic table bench.py reports as roofline.traffic.
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_k4_tf32 k4/tf32"""
import csv, io, json, os, subprocess, sys
rep, out, key = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
cols = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic"]
idx = [h.index(c) for c in cols if c in h]
with open(out + "_ncu_full_summary.csv", "w", newline="") as f:
    wr = csv.writer(f)
    wr.writerow([h[i] for i in idx]); wr.writerow([rows[1][i] for i in idx])
    for r in rows[2:]:
        wr.writerow([r[i] for i in idx])
names = {"tc_fwd_qmf_kernel": "tc_forward_qmf", "tc_bwd_qmf_kernel": "tc_backward_qmf", "modulate_stats_kernel": "modulate_stats",
         "modulate_apply_kernel": "modulate_apply", "narrow_kernel<0, 0": "narrow_step_jlogits", "narrow_kernel<1, 1": "narrow_forward_qmf",
         "narrow_kernel<1, 2": "narrow_backward_qmf", "pool_mean_kernel": "pool_mean", "pool_mean_bwd_kernel": "pool_mean_bwd", "rows_forward_vec_kernel<1": "rows_forward_qmf", "rows_forward_vec_kernel<0": "rows_forward_jlogits",
         "rows_backward_vec_kernel<1": "rows_backward_qmf", "rows_backward_vec_kernel<0": "rows_calibrated",
         "mid_kernel": "step_mid", "finalize_grads_kernel": "finalize_grads", "finalize_stats_kernel": "finalize_stats",
         "rows_forward_reg_kernel<1": "rows_forward_qmf", "rows_forward_reg_kernel<0": "rows_forward_jlogits",
         "rows_backward_reg_kernel<1": "rows_backward_qmf", "rows_backward_reg_kernel<0": "rows_calibrated"}
traffic = {}
unit = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
ki, ri, wi, gi = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("Grid Size")
for r in rows[2:]:
    n = r[ki]
    label = next((v for k, v in names.items() if k in n), None)
    if label is None and "tc_gemm_kernel" in n:
        order = os.environ.get("TC_ORDER", "tc_dfeat,tc_dweight,tc_logits").split(",")
        label = order[len([1 for k in traffic if k in order]) % len(order)]
    if label is None and "tc_heads_forward" in n:
        label = "tc_heads_forward"
    if label is None and "lf::" in n:
        import re
        m = re.search(r"lf::(\w+)", n)
        label = m.group(1) if m else None
    if label and label not in traffic:
        traffic[label] = int(float(r[ri]) * unit[rows[1][ri]] + float(r[wi]) * unit[rows[1][wi]])
path = os.path.join(os.path.dirname(out), "traffic.json")
allt = json.load(open(path)) if os.path.exists(path) else {}
allt[key] = traffic
json.dump(allt, open(path, "w"), indent=1, sort_keys=True)
print(traffic)
