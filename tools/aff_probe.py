"""Developer probe for ncu: ONE stand-alone MSM at 2^log_n with the pair tree on (group, levels from the command line).
Usage: ncu --set full -k regex:k_affine_level -c 2 python tools/aff_probe.py 2 5 [log_n]"""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
zk = importlib.import_module("zksnark-rs_b200")


def main():
    group, levels = int(sys.argv[1]), int(sys.argv[2])
    lg = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    os.environ["ZKB_AFF_G1" if group == 1 else "ZKB_AFF_G2"] = str(levels)
    n = 1 << lg
    ctx = zk.Context(0)
    rng = np.random.default_rng(1)
    k = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    k[:, 3] &= np.uint64((1 << 60) - 1)
    b = zk.Bases.generate(ctx, group, k)
    s = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    s[:, 3] &= np.uint64((1 << 60) - 1)
    print(zk.msm(ctx, b, s) is not None)


if __name__ == "__main__":
    main()
