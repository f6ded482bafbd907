"""The HBM-bound kernels alone (embedding gather, de-duplicated sparse Adagrad) at bench.py's roofline size.
    python scripts/hbm_microbench.py                 timed with CUDA events (bench.hbm_rooflines)
    python scripts/hbm_microbench.py --once          one launch of each, for `ncu --set full -k regex:"gather_concat|sparse_"`
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hm-retrieval-two-tower_b200"))

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--ids", type=int, default=1 << 20)
    ap.add_argument("--e", type=int, default=64)
    args = ap.parse_args()
    from pkg import _native as N

    lib = N.load()
    if args.once:
        bench.HBM_TIMING = (0, 1)
    for r in bench.hbm_rooflines(bench.peaks(), lib, b=args.ids, e=args.e):
        print(json.dumps(r))


if __name__ == "__main__":
    main()
