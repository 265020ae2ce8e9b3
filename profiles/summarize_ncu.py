"""Turns `ncu -i X.ncu-rep --page raw --csv` into the short per-kernel summary kept under
profiles/ (the .ncu-rep files themselves stay in gpurun_out/, which is scratch).
usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/NAME.json"""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_per_block",
    "launch__occupancy_limit_registers": "occ_limit_registers_blocks",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem_blocks",
    "launch__occupancy_limit_warps": "occ_limit_warps_blocks",
    "launch__waves_per_multiprocessor": "waves_per_sm",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct_of_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed": "fp64_pipe_pct_of_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__cycles_elapsed.max": "sm_cycles_elapsed_max",
    "smsp__cycles_active.avg": "smsp_cycles_active_avg",
}


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {}
        for i, h in enumerate(hdr):
            if h == "Kernel Name":
                d["kernel"] = vals[i].split("(")[0]
            if h in KEYS:
                d[KEYS[h]] = "%s %s" % (vals[i], units[i])
            if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct"):
                try:
                    v = float(vals[i])
                except ValueError:
                    continue
                if v >= 3.0:
                    d.setdefault("stall_pct_per_warp_active", {})[h.split("stalled_")[1].replace("_per_warp_active.pct", "")] = v
        res.append(d)
    json.dump({"source": rep, "launches": res}, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
