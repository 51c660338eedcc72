import os, sys, tempfile
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench, helpers as H
from tweeker_raytracer_b200 import host
args = bench.parse_args([])
tmp = tempfile.mkdtemp()
app = host.App(bench.system_file(tmp, args, 0), bench.scene_file(tmp, args))
app.render(1); app.synchronize()
app.restart(); app.render(1)
fr = app.frame_view()
print("N=1:", bench.parity_check(app, fr, 1, app.spp))
w, h = app.resolution
ref, sysd = H.oracle_scene(app), H.oracle_sys(app)
rows = np.arange(0, h, 64)
xy = np.array([(x, y) for y in rows for x in range(w)], dtype=np.uint32)
L0 = ref.path_radiance(sysd, app.info.miss, w, xy, 0).reshape(len(rows), w, 3)
got = np.asarray(fr).reshape(h, w, 4)[::64][..., :3]
d = np.abs(got - L0)
print("max diff vs L0", d.max(), "at", np.unravel_index(d.argmax(), d.shape), "got", got.reshape(-1,3)[d.reshape(-1,3).max(axis=1).argmax()], "want", L0.reshape(-1,3)[d.reshape(-1,3).max(axis=1).argmax()])
print("fraction of exactly equal values", float((got == L0).mean()))
app.close()
