"""Turn gpurun_out/ ncu artefacts into the text summaries committed under profiles/ (dev tool)."""
import collections, csv, subprocess, sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def launches(path, out, skip, count):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [r for r in list(csv.DictReader(lines))[skip:skip + count] if "scatt::" in r["Kernel Name"]]  # drop torch's flush / copy kernels
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

def raw_metrics(rep, out, want, idx=0):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    hdr, units, vals = r[0], r[1], r[2 + idx]
    got = {}
    with open(out, "w") as fh:
        fh.write(f"# ncu --set full --clock-control none, one launch (launch #{idx} of the report); source: {os.path.basename(rep)}\n")
        for h, u, v in zip(hdr, units, vals):
            if any(h == w or h.startswith(w) for w in want):
                fh.write(f"{h} [{u}] = {v}\n")
                got[h] = (u, v)
    print(open(out).read())
    return got

def full_step(rep, out, traffic_json=None):
    """Per-launch table of a whole-step `ncu --set full` report (every kernel of one eager step)."""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    def g(r, name, default="0"):
        i = col.get(name)
        return r[i] if i is not None and r[i] != "" else default
    def f(r, name):
        try:
            return float(g(r, name).replace(",", ""))
        except ValueError:
            return 0.0
    recs = []
    for r in rows[2:]:
        name = g(r, "Kernel Name").split("(")[0].replace("void ", "").replace("scatt::<unnamed>::", "").replace("scatt::(anonymous namespace)::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
        recs.append(dict(
            name=name, grid=g(r, "launch__grid_size"), regs=g(r, "launch__registers_per_thread"),
            ns=f(r, "gpu__time_duration.sum"), rd=f(r, "dram__bytes_read.sum"), wr=f(r, "dram__bytes_write.sum"),
            l2=f(r, "lts__t_bytes.sum"), tensor=f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            sm=f(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"), dram=f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            ru=units(rows[1], col, "dram__bytes_read.sum"), wu=units(rows[1], col, "dram__bytes_write.sum"), l2u=units(rows[1], col, "lts__t_bytes.sum"),
            tu=units(rows[1], col, "gpu__time_duration.sum")))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "usecond": 1e3, "nsecond": 1.0, "msecond": 1e6}
    for r in recs:
        r["ns"] *= scale.get(r["tu"], 1.0)
        r["rd"] *= scale.get(r["ru"], 1.0)
        r["wr"] *= scale.get(r["wu"], 1.0)
        r["l2"] *= scale.get(r["l2u"], 1.0)
    tot = sum(r["ns"] for r in recs)
    agg = collections.OrderedDict()
    for r in recs:
        a = agg.setdefault(r["name"], dict(n=0, ns=0.0, rd=0.0, wr=0.0, l2=0.0, tensor=0.0, sm=0.0))
        a["n"] += 1
        for k in ("ns", "rd", "wr", "l2"):
            a[k] += r[k]
        a["tensor"] += r["tensor"] * r["ns"]
        a["sm"] += r["sm"] * r["ns"]
    with open(out, "w") as fh:
        fh.write(f"# ncu --set full --clock-control none --profile-from-start off, one eager step ({len(recs)} launches, sum {tot/1e3:.1f} us; cold-cache, serialised)\n")
        fh.write(f"# source: {os.path.basename(rep)}; DRAM / L2 bytes are per launch averages; tensor% / sm% are time-weighted pct_of_peak_sustained_elapsed\n")
        fh.write(f"{'kernel':44s} {'n':>3s} {'total_us':>9s} {'avg_us':>8s} {'share':>6s} {'dram_rd_MB':>10s} {'dram_wr_MB':>10s} {'tensor%':>8s} {'sm%':>6s}\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
            fh.write(f"{k:44s} {a['n']:3d} {a['ns']/1e3:9.1f} {a['ns']/a['n']/1e3:8.2f} {100*a['ns']/tot:5.1f}% {a['rd']/a['n']/1e6:10.3f} {a['wr']/a['n']/1e6:10.3f} {a['tensor']/max(a['ns'],1):8.1f} {a['sm']/max(a['ns'],1):6.1f}\n")
        fh.write("\n# per launch, in stream order: kernel, grid, regs, us, dram_rd_MB, dram_wr_MB, tensor%, sm%, dram%\n")
        for r in recs:
            fh.write(f"{r['name']:44s} {r['grid']:>6s} {r['regs']:>4s} {r['ns']/1e3:8.2f} {r['rd']/1e6:8.3f} {r['wr']/1e6:8.3f} {r['tensor']:6.1f} {r['sm']:6.1f} {r['dram']:6.1f}\n")
    print(open(out).read().split("\n# per launch")[0])
    if traffic_json:
        tj = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch (average over the launches of that kernel in one eager step) from the ncu --set full capture summarised in " + os.path.basename(out) + " (cold L2 per replay pass)"}
        for k, a in agg.items():
            tj[k] = int((a["rd"] + a["wr"]) / a["n"])
        with open(traffic_json, "w") as fh:
            json.dump(tj, fh, indent=1)


def units(urow, col, name):
    i = col.get(name)
    return urow[i] if i is not None else ""


if __name__ == "__main__":
    WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
            "smsp__average_warps_issue_stalled", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum"]
    for a in sys.argv[1:]:
        kind, src, dst = a.split(":")[:3]
        if kind == "fullstep":
            full_step(src, dst, a.split(":")[3] if len(a.split(":")) > 3 else None)
        elif kind == "launches":
            skip, count = int(a.split(":")[3]), int(a.split(":")[4])
            launches(src, dst, skip, count)
        else:
            raw_metrics(src, dst, WANT, int(a.split(":")[3]) if len(a.split(":")) > 3 else 0)
