"""Condense an `ncu --set full` report into the per-kernel text summary kept under profiles/ and the
dram-traffic JSON bench.py reads (profiles/traffic_rN.json).
usage: python scripts/ncu_summary.py REPORT.ncu-rep OUT.txt [TRAFFIC.json]"""
import csv, io, json, subprocess, sys, collections

rep, out_txt = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "gpc__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
        "launch__cluster_size" if "launch__cluster_size" in hdr else "launch__grid_size"]
name_i = hdr.index("Kernel Name")
seen = collections.OrderedDict()
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    seen.setdefault(r[name_i], []).append(r)
traffic = {}
with open(out_txt, "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on   ({rep.split('/')[-1]}; B200; durations under ncu are\n"
            "# cold-cache and serialised).  One block per kernel; several launches of a kernel are averaged (n given).\n\n")
    for name, rs in seen.items():
        short = name.split("(")[0].replace("pcfa::", "").replace("void ", "")
        f.write(f"== {short}   (n={len(rs)})\n")
        for k in dict.fromkeys(keep):
            if k not in hdr:
                continue
            i = hdr.index(k)
            try:
                v = sum(float(r[i].replace(",", "")) for r in rs) / len(rs)
            except ValueError:
                continue
            f.write(f"   {k:78s} {units[i]:12s} {v:,.3f}\n")
        try:
            rd = hdr.index("dram__bytes_read.sum"); wr = hdr.index("dram__bytes_write.sum")
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            traffic[short] = int(sum(float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]] for r in rs) / len(rs))
        except Exception:
            pass
        f.write("\n")
print(open(out_txt).read()[:3000])
if len(sys.argv) > 3:
    k = traffic
    g = lambda *names: sum(k.get(n, 0) for n in names)
    ep = {"pcfa_corr_pyramid_forward": g("prep_targets_kernel", "corr_pyramid_tc2_kernel"),
          "pcfa_corr_pyramid_backward": g("bw_prep_kernel", "bw_unpool_kernel") + 2 * k.get("corr_pyramid_bwd_tc2_kernel", 0),
          "pcfa_corr_pyramid_backward_occ": g("bw_prep_kernel", "bw_unpool_kernel") + 2 * k.get("corr_pyramid_bwd_tc2_kernel", 0),
          "pcfa_corr_occupancy_mark": k.get("occ_mark_kernel", 0),
          "pcfa_corr_lookup_forward": k.get("corr_lookup_fwd_cl2_kernel<4, 4>", k.get("corr_lookup_fwd_kernel<4>", 0)),
          "pcfa_corr_lookup_backward": k.get("corr_lookup_bwd_cl2_kernel<4, 4>", k.get("corr_lookup_bwd_kernel<4>", 0))}
    json.dump({"source": out_txt + " (dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the captured launches)",
               "kernels": k, "entry_points": ep}, open(sys.argv[3], "w"), indent=1)
