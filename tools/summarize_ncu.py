"""Summarise ncu outputs under gpurun_out/ into small CSVs for profiles/ (run on the CPU box)."""
import collections, csv, subprocess, sys

def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in data:
        if len(r) > vi:
            name = r[ki].split("(")[0].replace("void ", "").replace("vrod::", "").replace("<unnamed>::", "")
            agg[name].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "avg_us", "share_pct"])
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            w.writerow([k, len(v), round(sum(v) / 1e3, 1), round(sum(v) / len(v) / 1e3, 2), round(sum(v) / tot * 100, 2)])
            print(f"{k:60s} n={len(v):4d} total={sum(v)/1e3:10.1f} us avg={sum(v)/len(v)/1e3:9.2f} us share={sum(v)/tot*100:5.1f}%")

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]

def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows) - 2)])
        for k in ["Kernel Name"] + KEEP:
            for i, h in enumerate(hdr):
                if h == k or h.endswith("." + k):
                    w.writerow([h, units[i]] + [r[i] for r in rows[2:]])
                    print(f"{h:95s} {units[i]:10s}", [r[i][:60] for r in rows[2:]])
                    break

if __name__ == "__main__":
    kind, src, dst = sys.argv[1:4]
    (launches if kind == "launches" else full)(src, dst)
