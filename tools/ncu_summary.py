#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the few numbers DESIGN.md / profiles/ quote.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [max_launches] > profiles/rNN_xxx.txt
"""
import csv
import subprocess
import sys

KEEP = ("Duration", "SM Frequency", "Registers Per Thread", "Achieved Occupancy", "Theoretical Occupancy",
        "Executed Ipc Active", "Issue Slots Busy", "Compute (SM) Throughput", "Memory Throughput", "DRAM Throughput",
        "L2 Cache Throughput", "L1/TEX Cache Throughput", "Executed Instructions", "Avg. Active Threads Per Warp",
        "Block Limit Registers", "Block Limit Shared Mem", "Warp Cycles Per Issued Instruction", "No Eligible",
        "Shared Memory Configuration Size", "Static Shared Memory Per Block")
RAW = ("dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "lts__t_bytes.sum",
       "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")


def main():
    rep = sys.argv[1]
    limit = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.DictReader(det.splitlines()))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units, body = rr[0], rr[1], rr[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# summary of {rep} (ncu --set full --clock-control none)")
    for lid in range(min(limit, len(body))):
        mine = [r for r in rows if r["ID"] == str(lid)]
        if not mine:
            continue
        name = mine[0]["Kernel Name"].replace("void <unnamed>::", "").split("(")[0]
        print(f"\n== launch {lid}: {name}  grid {mine[0]['Grid Size']} block {mine[0]['Block Size']}")
        seen = set()
        for r in mine:
            if r["Metric Name"] in KEEP and r["Metric Name"] not in seen:
                seen.add(r["Metric Name"])
                print(f"   {r['Metric Name']:38s} {r['Metric Value']:>14s} {r['Metric Unit']}")
        for m in RAW:
            if m in col:
                print(f"   {m:58s} {body[lid][col[m]]:>14s} {units[col[m]]}")
        rd, wr = (body[lid][col[m]] if m in col else "0" for m in RAW[:2])
        print(f"   stall reasons (top): " + ", ".join(
            f"{r['Metric Name']}={r['Metric Value']}" for r in mine if r["Section Name"] == "Warp State Statistics"
            and r["Metric Name"].startswith("Stall") ) [:400])


if __name__ == "__main__":
    main()
