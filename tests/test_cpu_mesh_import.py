"""`model assimp <file>` through the built-in Wavefront OBJ reader (host/MeshImport.cpp; reference: Assimp.cpp:47-319).
No GPU: the host library loads the scene, the oracle renders it."""
import os

import numpy as np

import helpers as H
from tweeker_raytracer_b200 import host

CUBE = """# unit cube, quads, two materials, no normals
mtllib cube.mtl
o cube
v -1 -1 -1
v  1 -1 -1
v  1  1 -1
v -1  1 -1
v -1 -1  1
v  1 -1  1
v  1  1  1
v -1  1  1
vt 0 0
vt 1 0
vt 1 1
vt 0 1
usemtl paint
f 1/1 4/4 3/3 2/2
f 5/1 6/2 7/3 8/4
f 1/1 2/2 6/3 5/4
usemtl metal
f 2/1 3/2 7/3 6/4
f 3/1 4/2 8/3 7/4
f -4/1 -8/2 -5/3 -1/4
o lid
usemtl nosuchmaterial
f 4//1 3//1 7//1
"""

MTL = """newmtl paint
Kd 0.25 0.5 0.75
newmtl metal
Ns 100
"""

SCENE = """albedo 0.5 0.5 0.5
material default brdf_diffuse
albedo 1 0 0
material paint brdf_diffuse
albedo 0.9 0.8 0.3
roughness 0.2 0.2
material metal brdf_ggx_smith
identity
push scale 10 1 10 model plane 1 1 1 default pop
push translate 0 1 0
model assimp {obj}
pop
push scale 0.5 0.5 0.5 translate 3 0.5 0
model assimp {obj}
pop
push
model assimp missing_file.obj
pop
"""


def _write(tmp_path, with_vn=False):
    d = str(tmp_path)
    obj = CUBE
    if with_vn:
        obj = obj.replace("vt 0 0\n", "vn 0 0 1\nvt 0 0\n", 1)
    else:
        obj = obj.replace("f 4//1 3//1 7//1", "f 4 3 7")
    with open(os.path.join(d, "cube.obj"), "w") as f:
        f.write(obj)
    with open(os.path.join(d, "cube.mtl"), "w") as f:
        f.write(MTL)
    scene = os.path.join(d, "scene_mesh.txt")
    with open(scene, "w") as f:
        f.write(SCENE.format(obj="cube.obj"))          # relative to the scene file
    return scene


def test_obj_import_structure(built, tmp_path):
    scene = _write(tmp_path)
    with host.App(H.write_system(tmp_path, "rtigo3_cornell_box", resolution="32 32", light=0, miss=1), scene, host_only=True) as app:
        # geometries: plane, then per (object, material): cube/paint, cube/metal, lid/nosuchmaterial; the second import instances them
        assert app.info.numGeometries == 4
        assert app.info.numInstances == 1 + 3 + 3
        names = {m: i for i, m in enumerate(["default", "paint", "metal"])}
        inst = [app.instance(i) for i in range(app.info.numInstances)]
        assert [g for _, g, _, _ in inst] == [0, 1, 2, 3, 1, 2, 3]
        assert [m for _, _, m, _ in inst] == [names["default"], names["paint"], names["metal"], names["default"]] + [names["paint"], names["metal"], names["default"]]
        # transforms: the model line's matrix reaches the meshes through two identity levels
        assert np.allclose(inst[1][0].reshape(3, 4), [[1, 0, 0, 0], [0, 1, 0, 1], [0, 0, 1, 0]])
        assert np.allclose(inst[4][0].reshape(3, 4), [[0.5, 0, 0, 3], [0, 0.5, 0, 0.5], [0, 0, 0.5, 0]])
        # Kd of the .mtl replaced the albedo of "paint" (Assimp.cpp:283-288); "metal" has no Kd and keeps the scene file's
        m = app.materials()
        assert np.allclose(m["albedo"][names["paint"]], [0.25, 0.5, 0.75]) and np.allclose(m["albedo"][names["metal"]], [0.9, 0.8, 0.3])
        # cube/paint: 3 quads -> 6 triangles, every corner its own vertex, fan order
        attrs, idx = app.geometry(1)
        assert idx.shape == (6, 3) and len(attrs) == 18 and np.array_equal(idx.ravel(), np.arange(18))
        assert np.array_equal(attrs["vertex"][:6], np.array([[-1, -1, -1], [-1, 1, -1], [1, 1, -1], [-1, -1, -1], [1, 1, -1], [1, -1, -1]], dtype=np.float32))
        assert np.array_equal(attrs["texcoord"][1], [0, 1, 0])
        # generated smooth normals: at a cube corner only the faces OF THIS MESH meeting there contribute (area weighted)
        corner = np.all(attrs["vertex"] == np.array([-1, -1, -1], dtype=np.float32), axis=1)
        n = attrs["normal"][corner]
        assert np.allclose(n, n[0]) and np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-6)
        assert np.allclose(n[0], np.array([0, -1, -1]) / np.sqrt(2), atol=1e-6)      # faces z=-1 and y=-1 of "paint"
        # tangents are orthogonal to the normals (calculateTangents)
        assert np.allclose(np.einsum("ij,ij->i", attrs["tangent"], attrs["normal"]), 0.0, atol=1e-6)
        assert np.allclose(np.linalg.norm(attrs["tangent"], axis=1), 1.0, atol=1e-6)
        # negative indices: the last face of cube/metal is x = -1
        a2, _ = app.geometry(2)
        assert np.all(a2["vertex"][-6:, 0] == -1.0)


def test_obj_normals_from_file_and_render(built, tmp_path):
    scene = _write(tmp_path, with_vn=True)
    with host.App(H.write_system(tmp_path, "rtigo3_cornell_box", resolution="40 30", light=2, miss=1, samplesSqrt=2, camera="0.7 0.55 50 9", center="0 1 0"), scene, host_only=True) as app:
        lid, _ = app.geometry(app.info.numGeometries - 1)      # geometry 0 is the area light's quad here
        assert np.array_equal(lid["normal"], np.tile(np.array([[0, 0, 1]], dtype=np.float32), (3, 1)))
        ref = H.oracle_scene(app)
        img = ref.render(H.oracle_sys(app), app.info.miss, 40, 30, iter_count=4).reshape(30, 40, 4)
        assert np.isfinite(img).all() and img[..., :3].mean() > 0.05
        # the cube is visible: rays through the image centre hit one of the imported geometries
        rays = ref.generate_primary(H.oracle_sys(app), 40, 30, 0)
        hits = ref.trace_closest(rays)
        hit_instances = set(hits["inst"][hits["inst"] != 0xffffffff].tolist())
        assert hit_instances & {2, 3, 4, 5, 6, 7}, hit_instances            # instance 0 = light quad, 1 = floor


def test_obj_reader_survives_bad_input(built, tmp_path):
    """Out-of-range and malformed face corners are skipped with a warning, lines / points / unknown statements are ignored,
    CRLF line ends and trailing blanks are tolerated, an empty file yields an empty model; nothing crashes."""
    d = str(tmp_path)
    with open(os.path.join(d, "bad.obj"), "w", newline="") as f:
        f.write("# comment\r\nv 0 0 0\r\nv 1 0 0\r\nv 0 1 0\r\nv 1 1 0 \r\n"
                "l 1 2\r\np 3\r\ns off\r\nbogus statement\r\n"
                "f 1 2 3\r\n"            # good
                "f 1 2 9\r\n"            # index out of range
                "f 1 2\r\n"              # too few corners
                "f a/b/c 2 3\r\n"        # garbage corner
                "f 2 4 3 \r\n")          # good, trailing blank
    with open(os.path.join(d, "empty.obj"), "w") as f:
        f.write("# nothing here\n")
    scene = os.path.join(d, "scene_bad.txt")
    with open(scene, "w") as f:
        f.write("material default brdf_diffuse\nidentity\npush\nmodel assimp bad.obj\npop\npush\nmodel assimp empty.obj\npop\n"
                "push\nmodel assimp model.fbx\npop\n")
    with open(os.path.join(d, "model.fbx"), "w") as f:
        f.write("not an obj")
    with host.App(H.write_system(tmp_path, "rtigo3_cornell_box", resolution="16 16", light=0, miss=1), scene, host_only=True) as app:
        assert app.info.numGeometries == 1 and app.info.numInstances == 1
        attrs, idx = app.geometry(0)
        assert idx.shape == (2, 3) and len(attrs) == 6
        assert np.array_equal(attrs["vertex"][3:], np.array([[1, 0, 0], [1, 1, 0], [0, 1, 0]], dtype=np.float32))
        assert np.allclose(attrs["normal"], [0, 0, 1])                     # generated: both triangles face +z
        img = H.oracle_scene(app).render(H.oracle_sys(app), app.info.miss, 16, 16, iter_count=1)
        assert np.isfinite(img).all()
