import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from oracle import oracle
from pytracer_b200 import _abi
from pytracer_b200.device import DeviceScene
from util import c1_params, demo_flat, golden
fs, cam = demo_flat()
g = golden("demo_c1_pathtracing_160x120.npz")
states = oracle.render(fs, c1_params(cam), want_states=True)["sample_states"]
sc = DeviceScene(fs)
res = {}
for prec in ("f64", "f32"):
    p = c1_params(cam, rng_mode=_abi.RT_RNG_REPLAY, variant="mega", precision=prec, out_f64=(prec=="f64"), hit_mode=_abi.RT_HIT_RAY_COUNT)
    rgb, cnt, st = sc.render(p, want_hit=True, replay_states=states)
    res[prec] = (rgb.astype(np.float64), cnt, st)
    print(prec, st["rays_closest"], cnt.sum())
d = res["f32"][1] - res["f64"][1]
print("pixels with different ray count:", (d != 0).sum(), "of", d.size, "sum diff", d.sum())
rows = np.nonzero((d != 0).any(axis=1))[0]
print("rows affected:", rows.min() if len(rows) else None, rows.max() if len(rows) else None)
hist = np.bincount(np.nonzero(d != 0)[0], minlength=120)
print("per-row count of differing pixels:", hist.tolist())
print("diff values sample:", np.unique(d[d != 0], return_counts=True))
hit = golden("demo_deterministic.npz")["hit_s0"]
for shp in (-1, 0, 1, 2):
    m = hit == shp
    print("primary-hit shape", shp, "pixels", m.sum(), "differing", (d[m] != 0).sum(), "sumdiff", d[m].sum())
bad = np.abs(res["f32"][0] - g["rgb"]).max(-1) > 1e-3 * np.maximum(np.abs(g["rgb"]).max(-1), 1e-3)
print("pixels with colour diff > 1e-3:", bad.sum())
