"""Turn a `gpurun_out/` ncu launch list + full-set report into the text summary committed under profiles/.
    python scripts/ncu_summary.py <launches.csv> <raw.csv from `ncu -i rep --page raw --csv`> > profiles/rNN_summary.md"""
import csv
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, durs = None, {}
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        key = (r[hdr.index("Kernel Name")][:72], r[hdr.index("Grid Size")], r[hdr.index("Block Size")])
        durs.setdefault(key, []).append(float(r[hdr.index("Metric Value")].replace(",", "")))
    tot = sum(sum(v) for v in durs.values())
    print("| kernel | grid | block | launches | total us | share | mean us |")
    print("|---|---|---|---|---|---|---|")
    for k, v in sorted(durs.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k[0]}` | {k[1]} | {k[2]} | {len(v)} | {sum(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% | {sum(v) / len(v) / 1e3:.2f} |")


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
            "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    for r in rows[2:]:
        print(f"\n**{r[hdr.index('Kernel Name')][:90]}**\n")
        print("| metric | value | unit |\n|---|---|---|")
        for w in want:
            if w in hdr:
                print(f"| {w} | {r[hdr.index(w)]} | {units[hdr.index(w)]} |")


if __name__ == "__main__":
    print("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`)\n")
    launches(sys.argv[1])
    print("\n## Dominant kernel (`ncu --set full --clock-control none --import-source on`)")
    raw(sys.argv[2])
