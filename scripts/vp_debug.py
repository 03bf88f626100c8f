import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petal_neighbors_b200 as pn
from petal_neighbors_b200 import synth
for n, d, bucket in ((300, 1, 8), (40, 1, 8), (20, 2, 8), (5000, 3, 32)):
    pts = synth.uniform(n, d, 200 + n + d, np.float32)
    h = pn.VantagePointTree.euclidean(pts, bucket_size=bucket, builder=pn.PN_BUILDER_HOST).layout()
    g = pn.VantagePointTree.euclidean(pts, bucket_size=bucket, builder=pn.PN_BUILDER_DEVICE).layout()
    print(n, d, bucket, "L", h["n_levels"], "radius eq", np.array_equal(h["node_radius"], g["node_radius"]), "center eq", np.array_equal(h["node_center"], g["node_center"]),
          "ids eq", np.array_equal(h["ids"], g["ids"]), "sorted ids eq", np.array_equal(np.sort(h["ids"]), np.sort(g["ids"])))
    bad = np.argwhere(h["ids"] != g["ids"]).ravel()
    print("  bad positions", bad[:20], "bucket_lo", h["bucket_lo"][:12], "bucket_hi", h["bucket_hi"][:12])
    if bad.size:
        print("  host", h["ids"][:40]); print("  dev ", g["ids"][:40])
        r = np.argwhere(h["node_radius"] != g["node_radius"]).ravel(); print("  radius mismatch nodes", r[:10])
