"""Turn gpurun_out/ ncu artefacts into the text summaries committed under profiles/ (dev tool)."""
import collections, csv, subprocess, sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def launches(path, out, skip, count):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))[skip:skip + count]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("void scatt::<unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += float(r["Metric Value"].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as fh:
        fh.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): one bench step = {len(rows)} launches, sum {tot/1e3:.1f} us\n")
        fh.write("# compare SHARES, not absolutes (B200_PROFILING.md)\n")
        fh.write(f"{'kernel':46s} {'launches':>8s} {'total_us':>10s} {'avg_us':>8s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"{k:46s} {v[0]:8d} {v[1]/1e3:10.1f} {v[1]/v[0]/1e3:8.2f} {100*v[1]/tot:6.1f}%\n")
        fh.write("\n# per launch, in stream order: id, kernel, grid, ns\n")
        for r in rows:
            fh.write(f"{r['ID']:>4s} {r['Kernel Name'].split('(')[0].replace('void scatt::<unnamed>::',''):46s} {r['Grid Size']:14s} {r['Metric Value']}\n")
    print(open(out).read().split("\n# per launch")[0])

def raw_metrics(rep, out, want):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    hdr, units, vals = r[0], r[1], r[2]
    got = {}
    with open(out, "w") as fh:
        fh.write(f"# ncu --set full --clock-control none, one launch; source: {os.path.basename(rep)}\n")
        for h, u, v in zip(hdr, units, vals):
            if any(h == w or h.startswith(w) for w in want):
                fh.write(f"{h} [{u}] = {v}\n")
                got[h] = (u, v)
    print(open(out).read())
    return got

if __name__ == "__main__":
    WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
            "smsp__average_warps_issue_stalled", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum"]
    for a in sys.argv[1:]:
        kind, src, dst = a.split(":")[:3]
        if kind == "launches":
            skip, count = int(a.split(":")[3]), int(a.split(":")[4])
            launches(src, dst, skip, count)
        else:
            raw_metrics(src, dst, WANT)
