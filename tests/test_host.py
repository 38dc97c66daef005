"""Host-side formats either side of the path: OBJ input, scene assembly, output.bmp."""
import numpy as np
import pytest

import ripoff_raytracer_b200 as rr
from oracle.pyoracle import Reference
from ripoff_raytracer_b200 import _abi, scenes

FIELDS = ["type", "ior", "color", "emissionColor", "emissionStrength", "reflectiveness", "specularProbability"]


def sorted_rows(t):
    return np.sort(np.ascontiguousarray(t).view(np.uint8).reshape(len(t), 96).view("V96").ravel())


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built")
def test_default_scene_equals_reference_scene(knight_obj):
    ref = Reference("strict")
    rt, rm, _ = ref.scene_default(knight_obj)
    s = rr.default_scene(knight_obj)
    t, m, r, sp = s.arrays()
    assert len(t) == len(rt) == 2222 and len(m) == len(rm) == 8 and len(sp) == 0
    for k in ["pos", "pitch", "yaw", "roll", "scale"]:
        assert np.array_equal(m[k], rm[k]), k
    for k in FIELDS:
        assert np.array_equal(m["material"][k], rm["material"][k]), k
    # the reference's SAH build permutes the OBJ triangles in place; same set, quads identical
    assert np.array_equal(sorted_rows(t[:2208]), sorted_rows(rt[:2208]))
    assert t[2208:].tobytes() == rt[2208:].tobytes()
    assert r["firstTriangle"].tolist() == [2208, 2210, 2212, 2214, 2216, 2218, 2220, 0]
    assert rr.default_camera(512, 512).tobytes() == ref.default_camera(512, 512).tobytes()


def test_default_scene_from_golden_obj(golden_small, tmp_path):
    """Scene assembly from the OBJ text stored with the golden vectors == the reference's arrays."""
    g = golden_small
    p = tmp_path / "k.obj"
    p.write_bytes(g["obj_text"].tobytes())
    t, m, r, _ = rr.default_scene(p).arrays()
    n_obj = len(t) - 14
    assert np.array_equal(sorted_rows(t[:n_obj]), sorted_rows(g["tris"][:n_obj]))
    assert t[n_obj:].tobytes() == g["tris"][n_obj:].tobytes()
    for k in FIELDS:
        assert np.array_equal(m["material"][k], g["meshes"]["material"][k]), k
    assert np.array_equal(r["firstTriangle"], g["ranges"]["firstTriangle"])
    assert np.array_equal(r["numTriangles"], g["ranges"]["numTriangles"])


def test_obj_loader_dialects_and_bad_faces(tmp_path):
    p = tmp_path / "m.obj"
    p.write_text(
        "# comment\n\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nvn 0 0 1\n"
        "f 1//1 2//1 3//1\n"        # v//vn
        "f 2/7/1 4/8/1 3/9/1\n"     # v/vt/vn
        "f 1//1 2//1 4//1 3//1\n"   # quad: 4th corner dropped (src/readobj.hpp:307-312)
        "f 1 2 3\n"                 # unsupported: skipped
        "f 1//1 2//1 9//1\n"        # out of bounds: skipped
        "f 1//1 2/7/1 3//1\n"       # mixed corner forms: neither sscanf pattern of the reference matches -> skipped
        "f 1/7/1 2//1 3/9/1\n"      # same
    )
    s = rr.Scene()
    mesh, rng = s.load_obj(p)
    assert int(rng["numTriangles"][0]) == 3 and int(rng["firstTriangle"][0]) == 0
    t = s.arrays()[0]
    assert t["posB"][1, :3].tolist() == [1.0, 1.0, 0.0] and t["posC"][2, :3].tolist() == [1.0, 1.0, 0.0]
    assert np.all(t["normalA"][:, :3] == [0, 0, 1])
    assert mesh["scale"][0] == 1.0 and mesh["material"]["type"][0] == _abi.MATERIAL_SOLID
    with pytest.raises(_abi.RRError):
        s.load_obj(tmp_path / "missing.obj")


def test_obj_roundtrip_is_bit_exact(tmp_path):
    v, n, f = scenes.displaced_icosphere(2, seed=7)
    p = tmp_path / "blob.obj"
    scenes.write_obj(p, v, n, f)
    s = rr.Scene()
    _, rng = s.load_obj(p)
    t = s.arrays()[0]
    assert int(rng["numTriangles"][0]) == len(f) == 320
    assert t.tobytes() == scenes.mesh_triangles(v, n, f).tobytes()


@pytest.mark.parametrize("shape", [(5, 7), (4, 8), (1, 1), (3, 2)])
def test_bmp_writer_is_byte_identical_to_reference(tmp_path, shape):
    H, W = shape
    rng = np.random.default_rng(W * 100 + H)
    rgba = rng.integers(0, 256, size=(H, W, 4), dtype=np.uint8)
    a = tmp_path / "a.bmp"
    rr.write_bmp(a, rgba)
    data = a.read_bytes()
    pad = (4 - (W * 3) % 4) % 4
    assert len(data) == 54 + (3 * W + pad) * H
    assert data[:2] == b"BM" and data[10] == 54 and data[14] == 40 and data[26] == 1 and data[28] == 24
    assert int.from_bytes(data[18:22], "little") == W and int.from_bytes(data[22:26], "little") == H
    row0 = data[54:54 + 3 * W]  # bottom row first, BGR
    assert row0 == rgba[H - 1, :, 2::-1].tobytes()
    if Reference.available():
        b = tmp_path / "b.bmp"
        Reference("strict").write_bmp(rgba, b)
        assert data == b.read_bytes()


def test_bmp_unwritable_path_reports_io_error(tmp_path):
    with pytest.raises(_abi.RRError) as e:
        rr.write_bmp(tmp_path / "no_such_dir" / "x.bmp", np.zeros((2, 2, 4), np.uint8))
    assert e.value.status == 8


def test_mesh_ranges_are_validated_without_a_device():
    # plan_segments runs before any CUDA call only when a context exists; here we only check the scene builder
    s = rr.Scene()
    bad = np.zeros(1, _abi.MESH_RANGE)
    bad["numTriangles"] = 5
    with pytest.raises(_abi.RRError) as e:
        s.add_mesh(np.zeros(1, _abi.MESH), bad)
    assert e.value.status == 6


def test_indexed_obj_loader_matches_the_reference_dialect(tmp_path):
    """rr_obj_load on the dialect the reference accepts == rr_scene_load_obj (== the reference loader, see above)."""
    v, n, f = scenes.displaced_icosphere(2, seed=7)
    p = tmp_path / "blob.obj"
    scenes.write_obj(p, v, n, f)
    pos, nrm, cor = rr.load_obj_indexed(p)
    assert len(cor) == len(f) == 320 and len(pos) == len(v) and len(nrm) == len(n)
    s = rr.Scene()
    s.load_obj(p)
    assert rr.triangles_from_indexed(pos, nrm, cor).tobytes() == s.arrays()[0].tobytes()


def test_indexed_obj_loader_lifts_the_reference_limits(tmp_path):
    p = tmp_path / "m.obj"
    p.write_text(
        "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0 0 2\nvn 0 0 1\n"
        "f 1//1 2//1 3//1 4//1\n"   # quad -> two triangles (the reference drops the 4th corner)
        "f 1 2 3\n"                 # no normals -> face normal (the reference skips the face)
        "f 1/5 2/6 5/7\n"           # v/vt
        "f -5//-1 -4//-1 -3//-1\n"  # relative indices
        "f 1//1 2//1 9//1\n"        # out of range: skipped
        "f 1//1 2//1\n"             # degenerate: skipped
    )
    pos, nrm, cor = rr.load_obj_indexed(p)
    assert len(pos) == 5 and len(cor) == 5
    assert cor[0].tolist() == [0, 1, 2, 0, 0, 0] and cor[1].tolist() == [0, 2, 3, 0, 0, 0]
    assert cor[2, :3].tolist() == [0, 1, 2] and np.allclose(nrm[cor[2, 3]], [0, 0, 1])      # face normal of a CCW triangle in z=0
    assert cor[3, :3].tolist() == [0, 1, 4] and np.allclose(nrm[cor[3, 3]], [0, -1, 0])
    assert cor[4].tolist() == [0, 1, 2, 0, 0, 0]
    with pytest.raises(_abi.RRError):
        rr.load_obj_indexed(tmp_path / "missing.obj")


def test_video_frame_setup_matches_reference(golden_video):
    """setupNextVideoFrame (src/image.hpp:385-390): yaw of the LAST mesh, float for float."""
    g = golden_video
    for count, idxs, want in [(int(g["video_count"]), range(int(g["video_count"])), g["video_yaw"]),
                              (1, range(4), g["video_yaw_count1"]), (360, (0, 1, 7, 180, 359), g["video_yaw_count360"])]:
        for idx, w in zip(idxs, want):
            m = np.zeros(3, _abi.MESH)
            m["yaw"] = 0.25
            rr.video_frame_setup(m, idx, count)
            assert m["yaw"][2].view(np.uint32) == np.float32(w).view(np.uint32), (count, idx)
            assert m["yaw"][0] == np.float32(0.25) and m["yaw"][1] == np.float32(0.25)  # only meshList.back() moves
    if Reference.available():
        ref = Reference("strict")
        ref.scene_set(np.zeros(0, _abi.TRIANGLE), np.zeros(2, _abi.MESH), np.zeros(0, _abi.GPU_NODE))
        for idx, count in [(3, 7), (59, 60), (0, 1)]:
            m = np.zeros(2, _abi.MESH)
            rr.video_frame_setup(m, idx, count)
            assert m["yaw"].tobytes() == ref.video_frame_setup(idx, count)["yaw"].tobytes()
    with pytest.raises(_abi.RRError):
        rr.video_frame_setup(np.zeros(1, _abi.MESH), 0, 0)


def test_video_frame_path():
    assert rr.video_frame_path("img", 1) == "img/output_1.bmp"      # src/main.cpp:701, render.sh:12
    assert rr.video_frame_path("/tmp/x", 120) == "/tmp/x/output_120.bmp"


def test_parallel_obj_parse_is_independent_of_the_chunking(tmp_path, monkeypatch):
    """The loader cuts the text into one chunk per thread; indices (1-based, relative, out of range) are resolved
    against the `v` / `vn` lines BEFORE a face in the whole file, whatever chunk they are in."""
    rng = np.random.default_rng(3)
    lines = ["# interleaved blocks of vertices, normals and faces"]
    nv = nn = 0
    for block in range(40):
        k = int(rng.integers(3, 9))
        for _ in range(k):
            x, y, z = rng.normal(size=3) * 10
            lines.append(f"v {x:.7g} {y:.7g} {z:.7g}" + ("\r" if block % 7 == 0 else ""))
        nv += k
        for _ in range(int(rng.integers(0, 3))):
            lines.append("vn 0 %.6f %.6f" % tuple(rng.normal(size=2)))
            nn += 1
        for _ in range(int(rng.integers(2, 7))):
            c = rng.integers(1, nv + 1, size=int(rng.integers(3, 6)))
            kind = int(rng.integers(0, 5))
            if kind == 0 or nn == 0:
                lines.append("f " + " ".join(str(i) for i in c))                         # no normals -> generated
            elif kind == 1:
                lines.append("f " + " ".join(f"{i}//{int(rng.integers(1, nn + 1))}" for i in c))
            elif kind == 2:
                lines.append("f " + " ".join(f"{i}/1/{int(rng.integers(1, nn + 1))}" for i in c))
            elif kind == 3:
                lines.append("f " + " ".join(f"{-int(rng.integers(1, nv + 1))}//-1" for _ in c))  # relative
            else:
                lines.append(f"f {nv + 5}//1 1//1 2//1")                                  # forward reference: skipped
    p = tmp_path / "mixed.obj"
    p.write_text("\n".join(lines))  # no trailing newline
    monkeypatch.setenv("RR_OBJ_MIN_CHUNK", "64")
    results = []
    for threads in (1, 2, 7, 16):
        monkeypatch.setenv("RR_OBJ_THREADS", str(threads))
        pos, nrm, cor = rr.load_obj_indexed(p)
        s = rr.Scene()
        s.load_obj(p)
        results.append((pos.tobytes(), nrm.tobytes(), cor.tobytes(), s.arrays()[0].tobytes()))
    assert len(cor) > 100 and all(r == results[0] for r in results[1:])
    # the forward references were skipped and every index is in range
    assert cor[:, :3].max() < len(pos) and cor[:, 3:].max() < len(nrm)


@pytest.mark.skipif(not __import__("pathlib").Path("/root/reference/src/readobj.hpp").exists(), reason="needs the reference headers")
def test_reference_binding_compiles_against_the_reference_headers(tmp_path):
    """include/reference_binding/image_b200.hpp compiled against the reference's OWN
    settings.hpp / readobj.hpp / math.hpp (read where they lie; the Khronos typedefs they need come from
    oracle/ref_shim/cl_host_types.hpp because the image has no CL headers) and linked with librr_b200.so: the
    wire-format static_asserts hold and every binding function instantiates.  The program also runs the part that
    needs no GPU: the mesh ranges recovered from a node list the reference's own loader + SplitBVH built."""
    import subprocess
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    src = tmp_path / "bind.cpp"
    src.write_text(r"""
#include "cl_host_types.hpp"
typedef cl_float3 float3_unused_;
#include "settings.hpp"
#include "readobj.hpp"
#include "reference_binding/image_b200.hpp"
int main(int argc, char** argv) {
  int n = -1;
  int rc = rr_device_count(&n);
  if (argc > 1) {  // scene through the reference's own loader, then the binding's upload path up to the device
    MeshInfo mesh = loadMeshFromOBJFile(argv[1]);
    meshList.emplace_back(mesh);
    std::cout << "triangles " << triangleList.size() << " nodes " << nodeList.size() << std::endl;
    if (rc == RR_OK && n > 0) {
      KernelContext k = generateKernelForDevices({0});
      generateBuffers(k, triangleList, meshList, nodeList);
      std::vector<unsigned char> pixels((size_t)WIDTH * HEIGHT * 4);
      CameraInformation cam{};
      cam.position = {CAMERA_START_X, CAMERA_START_Y, CAMERA_START_Z};
      cam.yaw = CAMERA_START_YAW; cam.fov = 90.0f; cam.aspectRatio = (float)WIDTH / (float)HEIGHT;
      compute(k, cam, pixels.data());
      release(k);
      std::cout << "rendered" << std::endl;
    }
  }
  std::cout << "binding ok, devices " << n << std::endl;
  return 0;
}
""")
    exe = tmp_path / "bind"
    csrc = root / "ripoff_raytracer_b200" / "csrc"
    cmd = ["g++", "-std=gnu++20", "-w", "-O1", "-o", str(exe), str(src), f"-I{root / 'oracle' / 'ref_shim'}", "-I/root/reference/src",
           f"-I{root / 'include'}", f"-L{csrc}", "-lrr_b200", f"-Wl,-rpath,{csrc}", "-pthread"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-4000:]
    obj = tmp_path / "k.obj"
    scenes.write_obj(obj, *scenes.uv_sphere(16, 8))
    p = subprocess.run([str(exe), str(obj)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "binding ok" in p.stdout, p.stdout + p.stderr
    assert "triangles 224" in p.stdout
