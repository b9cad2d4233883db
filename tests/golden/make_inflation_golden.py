#!/usr/bin/env python
"""Generates tests/golden/inflation_reference_dt.npz: seeded obstacle masks and the Euclidean distance maps the
REFERENCE's own code produces for them (nav2_costmap_2d/include/nav2_costmap_2d/distance_transform.hpp, compiled from
where it lies into oracle/_ref/libref_distance_transform.so by `make -C oracle ref`; needs /root/reference, so this
script only runs in the build container).  Committed so that the oracle (oracle/inflation_oracle.cpp) and the CUDA
kernel stay pinned to reference outputs on boxes where the reference is absent:
  tests/test_oracle_inflation.py::test_distance_map_matches_reference_golden   (CPU)
  tests/test_gpu_inflation.py::test_kernel_agrees_with_reference_golden        (GPU)
Run:  python tests/golden/make_inflation_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import inflation_binding as ob  # noqa: E402

CASES = [("sparse_97x131", (97, 131), 0.004, 1), ("dense_64x64", (64, 64), 0.15, 2), ("row_1x200", (1, 200), 0.03, 3),
         ("column_180x1", (180, 1), 0.05, 4), ("blocks_200x200", (200, 200), None, 5), ("single_33x65", (33, 65), -1, 6)]


def mask_for(shape, density, seed):
    rng = np.random.default_rng(seed)
    if density is None:  # a few axis-aligned blocks, like the benchmark maps
        m = np.zeros(shape, np.uint8)
        for _ in range(8):
            h, w = rng.integers(4, 13, size=2)
            y, x = rng.integers(0, shape[0] - h), rng.integers(0, shape[1] - w)
            m[y:y + h, x:x + w] = 1
        return m
    if density < 0:
        m = np.zeros(shape, np.uint8)
        m[shape[0] // 3, shape[1] // 2] = 1
        return m
    return (rng.random(shape) < density).astype(np.uint8)


def main():
    if ob.reference_distance_transform() is None:
        raise SystemExit("oracle/_ref/libref_distance_transform.so is missing: run `make -C oracle ref` where /root/reference exists")
    out = {}
    for name, shape, density, seed in CASES:
        mask = mask_for(shape, density, seed)
        out[f"{name}__mask"] = mask
        out[f"{name}__distance"] = ob.distance_map(mask, ref=True)
    path = os.path.join(HERE, "inflation_reference_dt.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
