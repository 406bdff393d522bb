"""Shared-memory data-pipe budget of each launch in an `ncu --set full` report (raw page CSV on stdin or as file):
LSU wavefronts (LDS / STS incl. bank conflicts), tensor-core operand wavefronts, and issue / tensor activity.
    ncu -i X.ncu-rep --page raw --csv | python tools/ncu_smem_pipe.py -"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "t_us"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed", "lds%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed", "sts%"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "tc_operand%"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_all%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue%"),
        ("sass__inst_executed_shared_loads", "LDS"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "lds_wf"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%")]


def main(path):
    f = sys.stdin if path == "-" else open(path)
    rows = list(csv.reader(f))
    hdr, data = rows[0], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(m, s) for m, s in COLS if m in idx]
    print("id  kernel".ljust(64) + "".join(s.rjust(13) for _, s in cols))
    for d in data:
        name = d[idx["Kernel Name"]].replace("void l2i::", "").replace("<unnamed>::", "").replace("unnamed>::", "")[:58]
        def fmt(v):
            try:
                x = float(v.replace(",", ""))
                return f"{x:13.1f}" if x < 1e6 else f"{x:13.3e}"
            except ValueError:
                return v.rjust(13)
        print(f"{d[idx['ID']]:>3} {name:60s}" + "".join(fmt(d[idx[m]]) for m, _ in cols))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "-")
