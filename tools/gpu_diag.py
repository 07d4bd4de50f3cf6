#!/usr/bin/env python3
"""Run every kernel parity check in its own subprocess (a trap in one kernel cannot poison the rest),
with a per-check timeout, and write a JSON report to gpurun_out/diag.json.

  python tools/gpu_diag.py [name-substring ...]
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_one(name):
    import torch
    from tests.gpu_checks import CHECKS
    t0 = time.time()
    try:
        res = CHECKS[name]()
        torch.cuda.synchronize()
        print("DIAG_RESULT " + json.dumps({"name": name, "ok": True, "res": res, "s": time.time() - t0}))
    except Exception as e:  # noqa
        print("DIAG_RESULT " + json.dumps({"name": name, "ok": False, "err": repr(e)[:600], "s": time.time() - t0}))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        return run_one(sys.argv[2])
    from tests.gpu_checks import CHECKS
    pats = sys.argv[1:]
    names = [n for n in CHECKS if not pats or any(p in n for p in pats)]
    report = []
    for n in names:
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", n], capture_output=True,
                               text=True, timeout=120, cwd=ROOT)
            line = [l for l in p.stdout.splitlines() if l.startswith("DIAG_RESULT ")]
            if line:
                r = json.loads(line[-1][len("DIAG_RESULT "):])
            else:
                r = {"name": n, "ok": False, "err": "no result; rc=%d stderr=%s" % (p.returncode, p.stderr[-600:])}
        except subprocess.TimeoutExpired:
            r = {"name": n, "ok": False, "err": "TIMEOUT"}
        report.append(r)
        print(("PASS " if r["ok"] else "FAIL ") + n + " " + json.dumps(r.get("res", r.get("err")))[:400], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
        json.dump(report, f, indent=1)
    bad = [r["name"] for r in report if not r["ok"]]
    print("SUMMARY: %d/%d passed; failed: %s" % (len(report) - len(bad), len(report), bad))


if __name__ == "__main__":
    main()
