"""Localise a slab-path discrepancy on a multi-GPU box: python scripts/slab_debug.py WORLD N DTYPE [NT]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _slab
import __graft_entry__ as g

if __name__ == "__main__":
    world, n, dtype = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    nt = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    res = _slab.run(world, "cuda", g.LIB, "debug_steps", n=n, dtype=dtype, nt=nt, timeout=600)
    for r, rr in enumerate(res):
        for s in rr["steps"]:
            print(f"rank {r} {s[0]:22s} rel={s[1]:.3e} maxabs={s[2]:.3e} at {s[3]} nbad={s[4]}")
