import numpy as np, vkhr_b200, oracle
from vkhr_b200 import capi
vox = vkhr_b200.Voxelizer(0)
P = oracle.port()
rng = np.random.default_rng(0)
W,H,D = 64,32,16
d = ((rng.random(W*H*D) < 0.15) * rng.integers(1,256,W*H*D)).astype(np.uint8)
try:
    g = vox.prefilter(d, W,H,D, ao=True, flags=capi.PREFILTER_GENERIC)["ao"]
    w = P.prefilter_ao(d, W,H,D)
    print("generic max rel", np.max(np.abs(g-w)/w))
except Exception as e: print("generic failed", e)
t = vox.prefilter(d, W,H,D, ao=True)["ao"]
print("tiled max rel", np.max(np.abs(t-w)/w))
