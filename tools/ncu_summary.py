"""Summarise an .ncu-rep (from `ncu --set full`) into a small markdown table for profiles/.
    python tools/ncu_summary.py gpurun_out/r01_full_L30.ncu-rep > profiles/r01_ncu_full_L30.md
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe % active"),
    ("sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed", "fp64 tensor %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global ld sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global ld requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global st sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "global st requests"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("nvlrx__bytes.sum", "nvlink rx"),
    ("nvltx__bytes.sum", "nvlink tx"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full summary of `%s`\n" % rep.split("/")[-1])
    print("Algorithmic bytes per launch of a gate pass = 32 B x 2^L (SURVEY 8d); `dram read+write` is the measured traffic.\n")
    for r in rows[2:]:
        print("## %s\n" % r[idx["Kernel Name"]])
        print("| metric | value | unit |\n|---|---|---|")
        for key, label in WANT:
            if key in idx and r[idx[key]] != "":
                print("| %s (`%s`) | %s | %s |" % (label, key, r[idx[key]], units[idx[key]]))
        try:
            rd = float(r[idx["dram__bytes_read.sum"]])
            wr = float(r[idx["dram__bytes_write.sum"]])
            u = units[idx["dram__bytes_read.sum"]]
            t = float(r[idx["gpu__time_duration.sum"]])
            tu = units[idx["gpu__time_duration.sum"]]
            print("| **traffic (read+write)** | %.3f | %s |" % (rd + wr, u))
            if u == "Gbyte" and tu == "ms":
                print("| **traffic / duration** | %.1f | GB/s (under ncu: cold, serialised) |" % ((rd + wr) / t * 1e3))
        except (KeyError, ValueError):
            pass
        print()


if __name__ == "__main__":
    main()
