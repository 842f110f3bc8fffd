"""Turn the ncu outputs under gpurun_out/ into the small text summaries committed under profiles/."""
import csv, io, subprocess, sys, collections
import numpy as np

def launch_table(path, out):
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    tot = collections.defaultdict(list)
    for r in rows:
        tot[r['Kernel Name'].split('(')[0].replace('void ', '')].append(float(r['Metric Value']) / 1e3)
    T = sum(sum(v) for v in tot.values())
    with open(out, "w") as f:
        f.write(f"# source: {path} (ncu --metrics gpu__time_duration.sum --clock-control none); times are cold-cache, serialised\n")
        f.write(f"# launches {len(rows)}  total {T/1e3:.2f} ms\n")
        f.write("%-34s %7s %10s %7s %9s %9s %9s\n" % ("kernel", "n", "sum_ms", "share", "median_us", "p90_us", "max_us"))
        for k, v in sorted(tot.items(), key=lambda kv: -sum(kv[1])):
            v = np.array(v)
            f.write("%-34s %7d %10.2f %6.1f%% %9.1f %9.1f %9.1f\n" % (k, len(v), v.sum() / 1e3, 100 * v.sum() / T, np.median(v), np.percentile(v, 90), v.max()))

def full_summary(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
            "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
    with open(out, "w") as f:
        f.write(f"# source: {rep} (ncu --set full --clock-control none --import-source on)\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            for k in keys:
                if k in d:
                    f.write("%-92s %s %s\n" % (k, d[k], units[hdr.index(k)]))
            f.write("\n")
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        idx = [i for i, r in enumerate(srows) if r and r[0] == "Kernel Name"]
        if idx:
            h = srows[idx[0] + 1]; body = srows[idx[0] + 2: (idx[1] if len(idx) > 1 else None)]
            si = h.index("# Samples")
            tot = sum(int(r[si]) for r in body if r[si].isdigit())
            by = collections.Counter()
            for r in body:
                if r[si].isdigit():
                    toks = r[1].split()
                    by[toks[1] if toks[0].startswith('@') else toks[0]] += int(r[si])
            f.write(f"# warp-stall samples by SASS opcode (total {tot}):\n")
            for k, v in by.most_common(12):
                f.write("  %-16s %6d %5.1f%%\n" % (k, v, 100 * v / max(tot, 1)))

if __name__ == "__main__":
    launch_table("gpurun_out/launches_r1_lap7_64_v3.csv", "profiles/r1_launches_lap7_64.txt")
    launch_table("gpurun_out/launches_r1_bench_lap7_128.csv", "profiles/r1_launches_bench_lap7_128.txt")
    full_summary("gpurun_out/prof_gemm128_r1_bigK.ncu-rep", "profiles/r1_ncu_full_gemm128_bigK_update.txt")
    full_summary("gpurun_out/prof_gemm128_r1_k256.ncu-rep", "profiles/r1_ncu_full_gemm128_k256_trailing.txt")
