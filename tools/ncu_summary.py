"""Summarise an ncu report (.ncu-rep) into markdown (+ optional traffic JSON for bench.py).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--traffic profiles/ncu_traffic.json] > profiles/x.md
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed"),
]


def main():
    rep = sys.argv[1]
    traffic_path = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    traffic = {}
    print(f"# ncu summary of `{rep.split('/')[-1]}`\n")
    print("Captured with `ncu --set full --clock-control none --import-source on`; per-launch values (cold-cache, serialised replays).\n")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        short = name.split("(")[0].replace("void ", "").replace("rag::", "")
        print(f"## `{short}`\n")
        print("| metric | value | unit |\n|---|---|---|")
        vals = {}
        for k, label in KEYS:
            if k in idx and r[idx[k]] != "":
                vals[k] = r[idx[k]]
                print(f"| {label} (`{k}`) | {r[idx[k]]} | {units[idx[k]]} |")
        st = []
        for h, i in idx.items():
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("\nwarp stall reasons per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in st[:8]) + "\n")

        def tobytes(k):
            if k not in vals:
                return None
            v, u = float(vals[k]), units[idx[k]].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

        rd, wr = tobytes("dram__bytes_read.sum"), tobytes("dram__bytes_write.sum")
        if rd is not None and wr is not None:
            key = short.split("<")[0]
            traffic.setdefault(key, int(rd + wr))
            print(f"DRAM traffic per launch: {(rd + wr) / 1e9:.4f} GB (read {rd / 1e9:.4f} + write {wr / 1e9:.4f})\n")
    if traffic_path:
        try:
            old = json.load(open(traffic_path))
        except Exception:
            old = {}
        old.update(traffic)
        json.dump(old, open(traffic_path, "w"), indent=1)


if __name__ == "__main__":
    main()
