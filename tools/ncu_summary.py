"""Summarises an `ncu --page raw --csv` export: one line per profiled launch with the roofline metrics."""
import csv
import sys

WANT = [("gpu__time_duration.sum", "t"), ("dram__bytes_read.sum", "dR"), ("dram__bytes_write.sum", "dW"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_tensor.sum", "tc_inst"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_bytes.sum", "L2B"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("sm__cycles_elapsed.avg.per_second", "clk")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(m, s) for m, s in WANT if m in idx]
    print("id  kernel" + " " * 44 + "grid  " + "  ".join(f"{s}[{units[idx[m]]}]" for m, s in cols))
    for d in data:
        name = d[idx["Kernel Name"]]
        name = name.replace("void l2i::", "").replace("<unnamed>::", "")[:48]
        print(f"{d[idx['ID']]:>3} {name:48s} {d[idx['Grid Size']]:>12} " + "  ".join(d[idx[m]] for m, _ in cols))


if __name__ == "__main__":
    main(sys.argv[1])
