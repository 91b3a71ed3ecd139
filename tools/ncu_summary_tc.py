"""Summarise the ncu report of one tensor-core projection step (tools/profile_tc.sh) into profiles/:
    python tools/ncu_summary_tc.py gpurun_out/prof_tc_<tag>.ncu-rep profiles/ncu_tc_path_<tag>.json profiles/traffic_tc.json
(reads the report with `ncu -i <rep> --page raw --csv`; run it where ncu is installed, no GPU needed)."""
import csv, io, json, subprocess, sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}

rep, out, traffic_out = sys.argv[1], sys.argv[2], sys.argv[3]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
head, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(head)}


def short(name):
    for k in ("tc_gemm_kernel", "tc_enc_kernel", "tc_head_kernel", "tc_split_weights_kernel"):
        if k in name:
            tail = name[name.index(k):]
            return tail[:tail.index("(")] if "(" in tail else tail
    return name


def num(r, k, table):
    return float(r[col[k]].replace(",", "")) * table.get(units[col[k]], 1.0)


res, tot_us, tot_rd, tot_wr, gemm_us = [], 0.0, 0.0, 0.0, 0.0
for r in data:
    d = {"kernel": short(r[col["Kernel Name"]])}
    for k in KEEP:
        if k in col:
            d[k] = f"{r[col[k]]} {units[col[k]]}".strip()
    res.append(d)
    us = num(r, "gpu__time_duration.sum", TIME)
    tot_us += us
    gemm_us += us if "tc_gemm" in d["kernel"] else 0.0
    tot_rd += num(r, "dram__bytes_read.sum", UNIT)
    tot_wr += num(r, "dram__bytes_write.sum", UNIT)
json.dump({"what": "every kernel of ONE tensor-core projection step over 65 536 poses (lrelu), `ncu --set full --clock-control none`, "
                   "in launch order; times are serialised, cold-cache ncu replays",
           "sum_us": tot_us, "gemm_share": gemm_us / tot_us, "kernels": res}, open(out, "w"), indent=1)
json.dump({"kernels": len(res), "gemm_share": gemm_us / tot_us, "step_us_under_ncu": tot_us, "dram_bytes_read_per_step": tot_rd, "dram_bytes_write_per_step": tot_wr,
           "dram_bytes_per_step": tot_rd + tot_wr, "step": "65 536 poses, 1 projection step on the tensor-core path (bench.py workload)",
           "note": "sum over the 15 launches of one step; the activations between the layer GEMMs (hi / lo planes, masks) are the "
                   "traffic -- algorithmic bytes of the step are 65 536 x 676",
           "source": "ncu --set full capture of tools/profile_tc.sh: " + rep}, open(traffic_out, "w"), indent=1)
print(f"{len(res)} launches, {tot_us:.1f} us, GEMM share {gemm_us / tot_us:.3f}, DRAM {1e-9 * (tot_rd + tot_wr):.2f} GB -> {out}, {traffic_out}")
