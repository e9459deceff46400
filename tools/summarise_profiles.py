#!/usr/bin/env python
"""Turn the raw ncu outputs of tools/capture_profiles.sh (gpurun_out/) into the small text summaries kept under profiles/:
  launch lists (ncu --metrics gpu__time_duration.sum --csv)  -> per-kernel launches / total / mean / share
  full captures (.ncu-rep, read with `ncu -i ... --page raw --csv`) -> the metrics the roofline discussion uses"""
import csv, io, os, subprocess, sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launch_summary(csv_path, out_path, title):
    rows = []
    with open(csv_path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    agg = OrderedDict()
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        name = r["Kernel Name"].split("(")[0][:48]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(out_path, "w") as o:
        o.write("# %s\n# ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold-cache launches: compare SHARES, not absolutes)\n" % title)
        o.write("%-48s %9s %10s %8s %6s\n" % ("kernel", "launches", "total us", "mean us", "share"))
        for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write("%-48s %9d %10.1f %8.2f %5.1f%%\n" % (name, n, us, us / n, 100 * us / tot))
        o.write("%-48s %9d %10.1f\n" % ("total", sum(a[0] for a in agg.values()), tot))


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def rep_summary(rep_path, out_path, title):
    out = subprocess.run(["ncu", "-i", rep_path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    with open(out_path, "w") as o:
        o.write("# %s\n# ncu --set full --clock-control none, %s; one column per captured launch\n" % (title, os.path.basename(rep_path)))
        for want in ["Kernel Name"] + KEYS:
            idx = [i for i, h in enumerate(hdr) if h == want or h.endswith("." + want)]
            if not idx:
                continue
            nonempty = [j for j in idx if any(r[j] for r in rows[2:])]
            i = (nonempty or idx)[0]
            o.write("%-100s %-10s %s\n" % (want, units[i], " | ".join(r[i] for r in rows[2:])))


if __name__ == "__main__":
    G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
    jobs = [("r02_launches.csv", "r02_launches_summary.txt", "bench.py --steps 30 (OPF-10k headline step), launch list"),
            ("r02_launches_c2.csv", "r02_launches_c2_dense_summary.txt", "config 2 (dense n=4096, m=2048): assemble / factorize / solve / solve_kkt / mul, launch list")]
    for src, dst, title in jobs:
        if os.path.exists(os.path.join(G, src)):
            launch_summary(os.path.join(G, src), os.path.join(P, dst), title)
    for src, dst, title in [("r02_prof_factor_dep.ncu-rep", "r02_prof_factor_dep_summary.txt", "k_factor_dep (numeric LDL^T of the OPF-10k tree in one launch)"),
                            ("r02_prof_update_bulk.ncu-rep", "r02_prof_update_bulk_summary.txt", "k_big_update_dyn_bulk (trailing update of the dense LDL^T, TMA bulk-copy staging + mbarrier ring + producer warp), N=4096"),
                            ("r02_prof_ozaki.ncu-rep", "r02_prof_ozaki_summary.txt", "ozk::k_ozaki_syrk v2 (tcgen05 int8 digits + TMA), K=2048 n=4096"),
                            ("r02_prof_dense_solve.ncu-rep", "r02_prof_dense_solve_summary.txt", "k_dense_solve_flow (dense triangular solves, one launch), N=4096")]:
        if os.path.exists(os.path.join(G, src)):
            rep_summary(os.path.join(G, src), os.path.join(P, dst), title)
    for f in ("r02_step_timeline.txt", "r02_profile_front.txt", "r02_c5md.json", "r02_sparse_timeline.txt", "r02_dense_timeline.txt", "r02_bench_1gpu.json",
              "r02_bench_reference.json"):
        if os.path.exists(os.path.join(G, f)):
            open(os.path.join(P, f), "w").write(open(os.path.join(G, f)).read())
