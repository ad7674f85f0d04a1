"""Turns `ncu -i capture.ncu-rep --page raw --csv` into (1) a slim per-launch CSV for profiles/ and (2) profiles/ncu_summary.json,
the per-kernel DRAM traffic / tensor-pipe activity that bench.py quotes in `roofline` (never typed into bench.py).

    ncu -i gpurun_out/r2_full.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summarize.py /tmp/raw.csv profiles/r2_ncu_full_layer2.csv "<command that was profiled>" """
import csv
import os, json, os, subprocess, sys

COLS = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_issued.avg.per_cycle_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "sm__cycles_elapsed.max",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
KEYS = [("cond_gemm_stage1", "tc_kernel<3, 128>"), ("cond_gemm", "tc_kernel<0, 256>"), ("kuf", "kuf_tc_kernel<256, "),
        ("dk_gemm", "dk_gemm_kernel<256, 2>"), ("dk_gemm_stage2", "dk_gemm_kernel<256, 1>"), ("dq_gemm", "xf_gemm_kernel<256>")]


def main():
    raw, out_csv, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    rows = list(csv.reader(open(raw, errors="replace")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    head, units, body = rows[hi], rows[hi + 1], rows[hi + 2:]
    idx = [head.index(c) for c in COLS if c in head]
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([head[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in body:
            if len(r) > max(idx):
                w.writerow([r[i] for i in idx])
    col = {c: head.index(c) for c in COLS if c in head}

    def val(r, c, scale_bytes=False):
        x = float(r[col[c]].replace(",", ""))
        return x * UNIT.get(units[col[c]], 1.0) if scale_bytes else x

    summ = {}
    for key, frag in KEYS:
        cand = [r for r in body if len(r) > max(idx) and frag in r[col["Kernel Name"]]]
        if not cand:
            continue
        r = max(cand, key=lambda r: val(r, "gpu__time_duration.sum"))      # the largest launch = conv layer 2
        summ[key] = {"kernel": r[col["Kernel Name"]][:80],
                     "dram_bytes": val(r, "dram__bytes_read.sum", True) + val(r, "dram__bytes_write.sum", True),
                     "tensor_pipe_active_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                     "ms_under_ncu": val(r, "gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0)}
    try:
        commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    except Exception:
        commit = ""
    commit = commit or os.environ.get("DCGP_TREE", "")      # (the GPU box has no .git: the caller passes the commit)
    summ["source"] = "%s (ncu --set full --clock-control none; %s; tree at %s)" % (out_csv, cmd, commit)
    json.dump(summ, open(os.path.join(os.path.dirname(out_csv) or ".", "ncu_summary.json"), "w"), indent=1)
    print(json.dumps(summ, indent=1))


if __name__ == "__main__":
    main()
