"""Writes the golden container read by test_dr_mvsnet in shims/dr_mvsnet.cpp (format documented there) from a
committed fixture (tests/golden/sample_*.npz), or - in the build container - from the reference's sample_inputs.pt.
    python tools/convert_sample_inputs.py tests/golden/sample_512x320.npz out/sample_inputs.bin [model tag]
"""
import struct
import sys

import numpy as np


def main(src, dst, tag="abl04"):
    g = np.load(src)
    V, H, W = g["bgr"].shape[:3]
    with open(dst, "wb") as f:
        f.write(b"TDMS0001")
        f.write(struct.pack("<iii", V, H, W))
        f.write(g["K3"].astype("<f4").tobytes())
        f.write(struct.pack("<fff", float(g["depth_min"]), float(g["depth_max"]), float(g["discard"])))
        f.write(g["c2w"].astype("<f4").tobytes())
        f.write(np.ascontiguousarray(g["bgr"]).tobytes())
        f.write(g[f"{tag}_stage3_depth"].astype("<f4").tobytes())
        f.write(g[f"{tag}_stage3_confidence"].astype("<f4").tobytes())


if __name__ == "__main__":
    main(*sys.argv[1:])
