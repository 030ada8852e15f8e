"""Run every GPU parity case in its own process (a CUDA fault poisons the context) and log the
numbers to gpurun_out/diag.jsonl.  Usage: python tools/gpu_diag.py [case ...]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_one(name):
    import gpu_cases
    t0 = time.time()
    m = gpu_cases.CASES[name]()
    ok = all(v[0] <= v[1] for v in m.values())
    print(json.dumps({"case": name, "ok": ok, "sec": round(time.time() - t0, 2),
                      "metrics": {k: [float(f"{v[0]:.3e}"), v[1]] for k, v in m.items()}}))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_one(sys.argv[2])
        return
    import gpu_cases
    names = sys.argv[1:] or list(gpu_cases.CASES)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", "diag.jsonl"), "a")
    for n in names:
        try:
            p = subprocess.run([sys.executable, __file__, "--one", n], capture_output=True, text=True, timeout=900)
            line = [l for l in p.stdout.splitlines() if l.startswith("{")]
            if p.returncode == 0 and line:
                rec = line[-1]
            else:
                rec = json.dumps({"case": n, "ok": False, "rc": p.returncode,
                                  "err": (p.stderr[-1500:] + p.stdout[-500:])})
        except subprocess.TimeoutExpired:
            rec = json.dumps({"case": n, "ok": False, "err": "timeout"})
        print(rec[:1800], flush=True)
        out.write(rec + "\n")
        out.flush()


if __name__ == "__main__":
    main()
