"""Run every kernel parity case in its own subprocess (a device trap in one case cannot hide the
others) and write gpurun_out/kernel_probe.json.  Usage: python tools/kernel_probe.py [pattern]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_one(name):
    import kernel_cases as kc
    from sscg_b200 import kernels as K
    err, scale, tol = kc.CASES[name]()
    print(json.dumps({"case": name, "err": err, "scale": scale, "tol": tol, "ok": bool(err <= tol),
                      "dev_error": K.device_error()}))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        run_one(sys.argv[2])
        return
    import kernel_cases as kc
    pat = sys.argv[1] if len(sys.argv) > 1 else ""
    out = []
    for name in kc.CASES:
        if pat and pat not in name:
            continue
        try:
            r = subprocess.run([sys.executable, __file__, "--case", name], capture_output=True, text=True, timeout=180)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if line:
                rec = json.loads(line[-1])
            else:
                rec = {"case": name, "ok": False, "rc": r.returncode, "stderr": r.stderr[-1500:]}
        except subprocess.TimeoutExpired:
            rec = {"case": name, "ok": False, "timeout": True}
        print(rec, flush=True)
        out.append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "kernel_probe.json"), "w") as f:
        json.dump(out, f, indent=1)
    bad = [r["case"] for r in out if not r.get("ok")]
    print(f"{len(out) - len(bad)}/{len(out)} cases ok; failing: {bad}")


if __name__ == "__main__":
    main()
