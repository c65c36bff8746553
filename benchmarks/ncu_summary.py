"""Summarise an `ncu --page raw --csv` export: one line per captured launch with the metrics DESIGN.md / bench.py quote.
usage: python benchmarks/ncu_summary.py gpurun_out/<tag>/c4_top_raw.csv [--traffic profiles/r02_traffic.json --config c4]
The rows key of the traffic file is the kernel's first grid-independent size argument as bench.py names it (rows of the launch);
ncu does not know it, so launches are matched to rows by order: the capture is the bench's own launch order per group."""
import csv
import json
import sys

WANT = {
    "gpu__time_duration.sum": "dur_us",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "inst",
    "sm__inst_executed_pipe_uniform.sum": "inst_uniform",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "ld_sectors",
}
UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def rows(path):
    with open(path, newline="") as f:
        r = list(csv.reader(f))
    hdr, units = r[0], r[1]
    out = []
    for line in r[2:]:
        d = {"kernel": line[hdr.index("Kernel Name")].split("(")[0]}
        for k, short in WANT.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(line[i].replace(",", ""))
                except ValueError:
                    continue
                v *= UNIT.get(units[i], 1.0)
                d[short] = v
        out.append(d)
    return out


def main():
    rs = rows(sys.argv[1])
    print("kernel,grid,block,regs,dur_us,dram_MB,l2_hit_pct,dram_pct,sm_pct,issue_pct,occ_pct,inst_M")
    for d in rs:
        print("%s,%d,%d,%d,%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,%.2f" % (
            d["kernel"], d.get("grid", 0), d.get("block", 0), d.get("regs", 0), d.get("dur_us", 0),
            (d.get("dram_rd", 0) + d.get("dram_wr", 0)) / 1e6, d.get("l2_hit_pct", 0), d.get("dram_pct", 0), d.get("sm_pct", 0),
            d.get("issue_pct", 0), d.get("occ_pct", 0), d.get("inst", 0) / 1e6))
    if "--traffic" in sys.argv:
        path = sys.argv[sys.argv.index("--traffic") + 1]
        cfg = sys.argv[sys.argv.index("--config") + 1]
        rows_by_kernel = json.loads(sys.argv[sys.argv.index("--rows") + 1])   # {"k_sage_mean": [rows of 1st, 2nd ... launch per group]}
        try:
            tj = json.load(open(path))
        except Exception:
            tj = {}
        ent = tj.setdefault(cfg, {}).setdefault("kernels", {})
        seen = {}
        for d in rs:
            nm = d["kernel"].split("<")[0].replace("void ", "").replace("eu::", "").strip()
            if nm not in rows_by_kernel:
                continue
            i = seen.get(nm, 0)
            seen[nm] = i + 1
            rr = rows_by_kernel[nm][i % len(rows_by_kernel[nm])]
            ent.setdefault(nm, {})[str(rr)] = {"dram_bytes_per_launch": int(d.get("dram_rd", 0) + d.get("dram_wr", 0)),
                                               "l2_hit_rate": round(d.get("l2_hit_pct", 0) / 100, 4), "ncu_duration_us": round(d.get("dur_us", 0), 1)}
        json.dump(tj, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
