"""Developer aid: compare two workspace dumps of tools/dev/casync_run (CASYNC_DUMP_WS) buffer by buffer."""
import sys
import numpy as np
BUFS = [("x1", 25600, 32), ("x2", 6400, 64), ("x3", 1600, 128), ("x4", 400, 256), ("cat", 100, 1024),
        ("d1t", 6400, 64), ("d2t", 1600, 128), ("d3t", 400, 256), ("d4t", 100, 512), ("aud_in", 1024, 32),
        ("a1", 1024, 64), ("a2", 1024, 128), ("a3", 256, 256), ("a4", 256, 256), ("a5", 100, 512),
        ("a6", 100, 512), ("fc1", 100, 1024), ("tx", 100, 1024), ("kall", 100, 256), ("vt", 2048, 128), ("p1q", 100, 576),
        ("att", 100, 512), ("ox0", 100, 1024), ("ox1", 100, 1024), ("ox2", 100, 1024),
        ("ox3", 100, 1024), ("kx", 100, 1024), ("f0", 100, 512), ("f1", 100, 512), ("f2", 100, 256),
        ("fuse", 100, 256), ("t_up1", 400, 128), ("up1", 400, 128), ("t_up2", 1600, 64), ("up2", 1600, 64),
        ("t_up3", 6400, 32), ("up3", 6400, 32), ("t_up4", 25600, 32), ("up4", 25600, 32)]
a = np.fromfile(sys.argv[1], dtype=np.uint16)
b = np.fromfile(sys.argv[2], dtype=np.uint16)
frames = int(sys.argv[3])
def f32(u):
    return (u.astype(np.uint32) << 16).view(np.float32)
off = 0
for name, rows, cols in BUFS:
    n = rows * cols * frames
    xa, xb = f32(a[off // 2: off // 2 + n]), f32(b[off // 2: off // 2 + n])
    d = np.abs(xa - xb)
    bad = np.nonzero(d > 0)[0]
    msg = ""
    if len(bad):
        r, c = bad // cols, bad % cols
        msg = " rows %d..%d (frames %d..%d) cols %d..%d first (%d,%d): %g vs %g" % (r.min(), r.max(), r.min() // rows, r.max() // rows,
                                                                                  c.min(), c.max(), r[0], c[0], xa[bad[0]], xb[bad[0]])
    if len(bad) and len(sys.argv) > 4 and sys.argv[4] == name:
        r, c = bad // cols, bad % cols
        uc, cc = np.unique(c, return_counts=True)
        ur, rc_ = np.unique(r, return_counts=True)
        print("  cols:", list(zip(uc.tolist(), cc.tolist()))[:80])
        print("  rows:", list(zip(ur.tolist(), rc_.tolist()))[:80])
    print("%-7s n_diff %8d  max %.4g%s" % (name, len(bad), d.max() if n else 0, msg))
    off += (n * 2 + 255) & ~255
