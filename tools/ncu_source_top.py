"""Hot spots of an `ncu --page source --csv` export (SASS view): executed-instruction histogram by opcode
and the instructions with the most stall samples."""
import collections
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != "Address"]
    tot_exec = sum(int(r[ix["Instructions Executed"]]) for r in data)
    tot_samp = sum(int(r[ix["# Samples"]]) for r in data)
    print(f"{len(data)} SASS instructions, {tot_exec} warp-instructions executed, {tot_samp} samples")
    by_op = collections.Counter()
    for r in data:
        op = r[ix["Source"]].split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
        by_op[op.split(".")[0]] += int(r[ix["Instructions Executed"]])
    print("executed by opcode:", ", ".join(f"{k} {100 * v / tot_exec:.1f}%" for k, v in by_op.most_common(18)))
    print("top stall-sample instructions:")
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
        st = sorted(((int(r[ix[c]]), c) for c in stall_cols), reverse=True)[:2]
        print(f"  {int(r[ix['# Samples']]):7d} {100 * int(r[ix['# Samples']]) / tot_samp:5.1f}%  exec {int(r[ix['Instructions Executed']]):9d}  "
              f"{r[ix['Source']].strip()[:70]:70s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
