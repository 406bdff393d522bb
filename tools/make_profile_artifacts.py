"""Turns one round's gpurun_out captures into the tracked artefacts under profiles/:
   python tools/make_profile_artifacts.py <tag>      (expects gpurun_out/{prof_<tag>_raw.csv, launches_<tag>.csv, layers_<tag>.json, bench_<tag>.log})"""
import collections
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(tag, n_forward_launches=22, batch=32):  # 17 conv + 5 FIR launches of one forward (ncu -k regex:"conv_tc|fir_tma")
    go, pr = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    rows = list(csv.reader(open(os.path.join(go, f"prof_{tag}_raw.csv"))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    tscale = {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}
    conv = fir = 0.0
    per = []
    for d in data[:n_forward_launches]:
        name = d[ix["Kernel Name"]]
        b = sum(float(d[ix[m]]) * scale[units[ix[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        if "fir_tma" in name:
            fir += b
        else:
            conv += b
        per.append({"id": int(d[ix["ID"]]), "kernel": name.replace("void l2i::<unnamed>::", "")[:60],
                    "ms": float(d[ix["gpu__time_duration.sum"]]) * tscale.get(units[ix["gpu__time_duration.sum"]], 1.0), "dram_bytes": b,
                    "tensor_pct": float(d[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])})
    json.dump({"source": f"profiles/{tag}_ncu_full_summary.txt (ncu --set full --clock-control none -k regex:conv_tc|fir_tma, bench.py --steps 1 "
                         f"--warmup 0, B={batch}, 1024px, first forward: the {n_forward_launches} conv / FIR launches)",
               "batch": batch, "conv_dram_bytes_per_image": conv / batch, "fir_dram_bytes_per_image": fir / batch, "launches": per},
              open(os.path.join(pr, "ncu_traffic.json"), "w"), indent=1)
    lines = [l for l in open(os.path.join(go, f"launches_{tag}.csv")) if not l.startswith("==")]
    r = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for row in r:
        a = agg.setdefault(row["Kernel Name"][:70], [0, 0.0])
        a[0] += 1
        a[1] += float(row["Metric Value"].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    out = [f"{len(r)} launches (bench.py --steps 2 --warmup 1 under ncu --metrics gpu__time_duration.sum), total {tot / 1e6:.2f} ms "
           "(cold-cache, serialised: compare SHARES)"]
    out += [f"{a[0]:4d} {a[1] / 1e6:9.3f} ms {100 * a[1] / tot:5.1f}%  {k}" for k, a in sorted(agg.items(), key=lambda x: -x[1][1])]
    open(os.path.join(pr, f"{tag}_launch_shares.txt"), "w").write("\n".join(out) + "\n")
    shutil.copy(os.path.join(go, f"launches_{tag}.csv"), os.path.join(pr, f"{tag}_launches_bench_b{batch}_1024.csv"))
    shutil.copy(os.path.join(go, f"layers_{tag}.json"), os.path.join(pr, f"{tag}_layers_b{batch}_1024.json"))
    shutil.copy(os.path.join(go, f"bench_{tag}.log"), os.path.join(pr, f"{tag}_bench.json"))
    print(f"conv DRAM {conv / batch / 1e6:.1f} MB/image, fir {fir / batch / 1e6:.1f} MB/image; {len(r)} launches listed")


if __name__ == "__main__":
    main(sys.argv[1])
