"""GPU diagnostic for the row-shift descriptor semantics: runs the 7x7 cases with both base_offset modes."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import kernel_cases as kc
from sscg_b200 import _lib as L
for mode in (1, 2):
    for name, fn in (("fwd21", lambda m: kc.case_conv_fwd(N=2, H=12, W=200, Cin=64, Cout=21, k=7, pad=3, bias=True, shift=m)),
                     ("fwd_oob", lambda m: kc.case_conv_fwd(N=2, H=10, W=70, Cin=64, Cout=21, k=7, pad=3, reflect=False, explicit=False, shift=m)),
                     ("dgrad", lambda m: kc.case_conv_dgrad(N=2, H=12, W=150, Cin=64, Cout=21, k=7, pad=3, shift=m))):
        try:
            err, scale, tol = fn(mode)
            print(json.dumps({"mode": mode, "case": name, "err": err, "scale": scale, "ok": err <= tol}), flush=True)
        except Exception as e:
            print(json.dumps({"mode": mode, "case": name, "error": repr(e)[:300]}), flush=True)
