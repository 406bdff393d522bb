"""A/B of differently compiled builds of the library on the bench workload:
   python tools/ab_libs.py [--steps N] name=path.so ...   (each in its own process via L2I_LIB; prints images/s and the per-layer table)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(argv):
    steps = 40
    if argv and argv[0] == "--steps":
        steps = int(argv[1]); argv = argv[2:]
    rows = {}
    for spec in argv:
        name, path = spec.split("=", 1)
        env = dict(os.environ, L2I_LIB=os.path.abspath(path))
        pj = os.path.join(ROOT, "gpurun_out", f"ab_{name}.json")
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(steps), "--warmup", "5", "--no-cpu-baseline",
                              "--profile-json", pj], env=env, capture_output=True, text=True, cwd=ROOT)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(name, "FAILED", out.stderr[-400:]); continue
        d = json.loads(line[-1])
        layers = json.load(open(pj)) if os.path.exists(pj) else {}
        rows[name] = (d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"], layers)
        print(f"{name:10s} value {d['value']:8.1f}  e2e {d['e2e']['value']:8.1f}  sm {d['clocks']['sm_mhz']}", flush=True)
    names = list(rows)
    if not names:
        return
    def seg(l):
        return {k: v["ms"] for k, v in l.get("segments", {}).items()}
    tabs = {n: seg(rows[n][3]) for n in names}
    keys = [k for k in tabs[names[0]] if tabs[names[0]][k] > 0.1]
    print("layer".ljust(28) + "".join(n.rjust(10) for n in names))
    for k in keys:
        print(k.ljust(28) + "".join(f"{tabs[n].get(k, float('nan')):10.3f}" for n in names))


if __name__ == "__main__":
    main(sys.argv[1:])
