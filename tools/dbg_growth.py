"""Development aid: error growth of the throughput build against the oracle, v2 and v3 kernels side by side.
usage: python tools/dbg_growth.py <case> [key=value ...] (run on the GPU box)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from gdtk_b200 import _abi, cases
from util import run_case, max_rel_diff

def main():
    name = sys.argv[1]
    kw = {}
    for a in sys.argv[2:]:
        k, v = a.split("=")
        try: v = int(v)
        except ValueError:
            try: v = float(v)
            except ValueError: pass
        kw[k] = v
    steps = kw.pop("steps", "1,2,5,10,20")
    steps = [int(s) for s in str(steps).split(",")]
    product = _abi.load_library()
    oracle = _abi.load_library(os.path.join(ROOT, "oracle", "_build", "liboracle.so"), "orc_")
    fac = getattr(cases, name)
    for n in steps:
        _, Uo, _ = run_case(fac, oracle, n, **kw)
        row = [f"steps {n:4d}"]
        for label, knob, strict in (("v3 strict", 0, True), ("v2 fast", 2, False), ("v3 fast", 0, False)):
            s, U, _ = run_case(fac, product, n, strict=strict, force_generic_kernel=knob, **kw)
            row.append(f"{label}: {max_rel_diff(U, Uo):.3e}")
            s.close()
        print("  ".join(row), flush=True)

main()
