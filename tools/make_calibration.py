#!/usr/bin/env python3
"""Pack the grid-camera calibration the reference loads at run time into one small .npz.

Source (data, read-only, only present in the build container):
  /root/reference/data/srv1/grid_calib/192.168.0.114/{phase2_160x120.csv, gripperBody_160x120.png,
  gripperHole_160x120.png}  -- the directory and resolution hard-wired in src/sensors/GridCamera.java:24-25.
Semantics follow src/sensors/GridCalibration.java:53-114: header line skipped, '#' lines skipped,
Xr/Yr parsed as float (+ CAMERA_X/Y = 0), masks = blue channel != 0.

Output: puckswarm_b200/data/calib_114_160x120.npz with
  xy    uint16 [n,2]  pixel (x, y) in CSV order
  XrYr  float32 [n,2]
  flags uint8 [n]     bit0 gripperBody, bit1 gripperHole
  size  int32 [2]     image width, height
"""
import os
import sys
import numpy as np
from PIL import Image

SRC = "/root/reference/data/srv1/grid_calib/192.168.0.114"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "puckswarm_b200", "data", "calib_114_160x120.npz")


def main():
    w, h = 160, 120
    rows = []
    with open(os.path.join(SRC, "phase2_160x120.csv")) as f:
        lines = f.read().splitlines()
    for s in lines[1:]:
        if not s or s[0] == "#":
            continue
        a = [t.strip() for t in s.split(",")]
        rows.append((int(a[0]), int(a[1]), np.float32(a[2]), np.float32(a[3])))
    body = np.asarray(Image.open(os.path.join(SRC, "gripperBody_160x120.png")).convert("RGB"))
    hole = np.asarray(Image.open(os.path.join(SRC, "gripperHole_160x120.png")).convert("RGB"))
    # a later CSV row for the same pixel replaces an earlier one (calibData[xp][yp] = new ...)
    last = {}
    for x, y, Xr, Yr in rows:
        last[(x, y)] = (Xr, Yr)
    keys = list(last.keys())
    xy = np.array(keys, dtype=np.uint16)
    XrYr = np.array([last[k] for k in keys], dtype=np.float32)
    flags = np.zeros(len(keys), dtype=np.uint8)
    for i, (x, y) in enumerate(keys):
        if body[y, x, 2] != 0:
            flags[i] |= 1
        if hole[y, x, 2] != 0:
            flags[i] |= 2
    np.savez_compressed(OUT, xy=xy, XrYr=XrYr, flags=flags, size=np.array([w, h], dtype=np.int32))
    print("rows", len(keys), "gripperBody", int((flags & 1).sum()), "gripperHole", int((flags & 2 > 0).sum()),
          "Xr", XrYr[:, 0].min(), XrYr[:, 0].max(), "Yr", XrYr[:, 1].min(), XrYr[:, 1].max())


if __name__ == "__main__":
    sys.exit(main())
