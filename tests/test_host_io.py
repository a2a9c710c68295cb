"""On-disk formats either side of the alignment path: OpenCV-XML frame dumps, pose text files, TUM
ground-truth trajectories.  The XML reader / writer is pinned against cv2.FileStorage (the library the reference calls,
src/SolveDVO.cpp:156, src/camTopic2PublisherPyD.cpp:325); the rest against direct numpy restatements."""
import numpy as np
import pytest

import host_lib as HL
import oracle_lib as O


def _levels(seed, W=32, H=24, L=4):
    rng = np.random.default_rng(seed)
    mono = [rng.integers(0, 256, (H >> l, W >> l), dtype=np.uint8) for l in range(L)]
    depth = [rng.integers(0, 65536, (H >> l, W >> l), dtype=np.uint16) for l in range(L)]
    return mono, depth


def test_frame_xml_round_trip(tmp_path):
    mono, depth = _levels(1)
    f = tmp_path / "framemono_0001.xml"
    assert HL.store_frame_xml(f, mono, depth) == 0
    rc, m2, d2 = HL.load_frame_xml(f, 4)
    assert rc == 0
    for l in range(4):
        assert np.array_equal(m2[l], mono[l]) and np.array_equal(d2[l], depth[l])
    assert HL.load_frame_xml(tmp_path / "missing.xml", 4)[0] == -1          # loadFromFile returns false (src/SolveDVO.cpp:158-162)
    assert HL.load_frame_xml(f, 5)[0] == -1                                  # a level the dump does not hold


def test_frame_xml_is_opencv_filestorage_compatible(tmp_path):
    cv2 = pytest.importorskip("cv2")
    mono, depth = _levels(2)
    # (a) cv2 reads what the writer produced
    f = tmp_path / "ours.xml"
    assert HL.store_frame_xml(f, mono, depth) == 0
    fs = cv2.FileStorage(str(f), cv2.FILE_STORAGE_READ)
    assert fs.isOpened()
    for l in range(4):
        assert np.array_equal(fs.getNode(f"mono_{l}").mat(), mono[l])
        assert np.array_equal(fs.getNode(f"depth_{l}").mat(), depth[l])
    fs.release()
    # (b) the reader parses what cv2 writes, including the node types the path does not use itself
    g = tmp_path / "cv.xml"
    fs = cv2.FileStorage(str(g), cv2.FILE_STORAGE_WRITE)
    rng = np.random.default_rng(3)
    extra = {"f": rng.standard_normal((5, 7)).astype(np.float32), "d": rng.standard_normal((3, 4)), "c3": rng.integers(0, 256, (4, 6, 3), dtype=np.uint8),
             "i": rng.integers(-2**31, 2**31 - 1, (3, 3), dtype=np.int32), "s": rng.integers(-2**15, 2**15 - 1, (2, 9), dtype=np.int16)}
    for l in range(4):
        fs.write(f"mono_{l}", mono[l]); fs.write(f"depth_{l}", depth[l])
    for k, v in extra.items():
        fs.write(k, v)
    fs.release()
    rc, m2, d2 = HL.load_frame_xml(g, 4)
    assert rc == 0
    for l in range(4):
        assert np.array_equal(m2[l], mono[l]) and np.array_equal(d2[l], depth[l])
    for k, v in extra.items():
        rc, got, _ = HL.read_xml_matrix(g, k)
        assert rc == 0 and got.shape == v.shape
        assert np.array_equal(got, v.astype(np.float64)), k                 # cv2 prints enough digits to round-trip fp32 / fp64
    assert HL.read_xml_matrix(g, "nope")[0] == -2


def test_golden_cv2_frame_dump():
    """A dump written by cv2.FileStorage in the build container (tests/golden/make_golden.py) -- no cv2 needed to read it."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "framemono_0000.xml")
    want = np.load(os.path.join(os.path.dirname(__file__), "golden", "framemono_0000.npz"))
    rc, m, d = HL.load_frame_xml(path, 4)
    assert rc == 0
    for l in range(4):
        assert np.array_equal(m[l], want[f"mono_{l}"]) and np.array_equal(d[l], want[f"depth_{l}"])
        # the publisher's rule: level l is the NEAREST subsample of the stored level 0 (src/camTopic2PublisherPyD.cpp:338-348)
        assert np.array_equal(m[l], m[0][:: 1 << l, :: 1 << l])


def test_pose_text_file(tmp_path):
    rng = np.random.default_rng(4)
    q = rng.standard_normal((20, 7))
    f = tmp_path / "estPoses.txt"
    assert HL.write_pose_file(f, q) == 0
    lines = open(f).read().splitlines()
    assert len(lines) == 20
    # SolveDVO::printPose streams doubles with the default ostream format (6 significant digits), "qx qy qz qw tx ty tz"
    assert lines[0] == " ".join("%g" % v for v in q[0])
    back = HL.read_pose_file(f)
    assert back.shape == (20, 7) and np.allclose(back, q, rtol=1e-5, atol=1e-6)


def _quat_to_R(q):      # (x, y, z, w) -> R, fp32 like Eigen::Quaternionf::toRotationMatrix
    x, y, z, w = [np.float32(v) for v in q]
    two = np.float32(2)
    tx, ty, tz = two * x, two * y, two * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz, tyy, tyz, tzz = tx * x, ty * x, tz * x, ty * y, tz * y, tz * z
    one = np.float32(1)
    return np.array([[one - (tyy + tzz), txy - twz, txz + twy], [txy + twz, one - (txx + tzz), tyz - twx], [txz - twy, tyz + twx, one - (txx + tyy)]], np.float32)


def test_tum_ground_truth_loader(tmp_path):
    rng = np.random.default_rng(5)
    n = 30
    q = rng.standard_normal((n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    t = rng.standard_normal((n, 3))
    f = tmp_path / "groundtruth.txt"
    with open(f, "w") as fh:
        fh.write("# ground truth trajectory\n# file: 'x.bag'\n   # timestamp tx ty tz qx qy qz qw\n")
        for i in range(n):
            # (TUM files carry 4 decimals; 8 here so that the quaternions stay unit to fp32 and R can be compared directly)
            fh.write(f"{1305031098.0 + i * 0.01:.4f} {t[i,0]:.4f} {t[i,1]:.4f} {t[i,2]:.4f} {q[i,0]:.8f} {q[i,1]:.8f} {q[i,2]:.8f} {q[i,3]:.8f}\n")
    skip = 4
    got = HL.load_gt_path(f, skip=skip)
    # the skip loop consumes skip + 1 data lines (src/loadGTPath.cpp:110-120), the next one becomes the world frame
    first = skip + 1
    assert got.shape == (n - first, 7)
    rows = [[np.float32(float(s)) for s in ln.split()] for ln in open(f).read().splitlines() if not ln.strip().startswith("#")]
    Rf = _quat_to_R(rows[first][4:8]); Tf = np.array(rows[first][1:4], np.float32)
    for k in range(first, n):
        Rc = _quat_to_R(rows[k][4:8]); Tc = np.array(rows[k][1:4], np.float32)
        Tu = Rf.T @ (Tc - Tf); Ru = Rf.T @ Rc
        g = got[k - first]
        assert np.allclose(g[4:7], Tu, atol=1e-5)
        Rg = _quat_to_R(g[0:4])
        assert np.allclose(Rg, Ru, atol=2e-5)
    assert np.allclose(got[0, 4:7], 0, atol=1e-7) and abs(abs(got[0, 3]) - 1) < 1e-5     # the first pose kept is the identity
    assert HL.load_gt_path(f, skip=-1).shape == (n, 7)                                       # no skipping
    assert HL.load_gt_path(f, skip=350).shape == (0, 7)                                      # the reference's hard-coded 350 on a short file
    assert HL.load_gt_path(tmp_path / "nope.txt") is None


@pytest.mark.gpu
def test_solvedvo_replays_frame_dumps(tmp_path):
    """SolveDVO::loadFromFile -> setRcvdFrameAs{Ref,Now}Frame -> runIterations on two XML dumps equals the oracle on the same images."""
    d = O.synth_batch(300, 1)
    W, H, L = 640, 480, 4
    iters = (20, 20, 20, 20)
    paths = []
    for name, gray, depth in (("ref", d["ref_gray"][0], d["ref_depth"][0]), ("now", d["now_gray"][0], d["now_depth"][0] if "now_depth" in d else d["ref_depth"][0])):
        mono = [gray[:: 1 << l, :: 1 << l] for l in range(L)]
        dep = [depth[:: 1 << l, :: 1 << l] for l in range(L)]
        p = tmp_path / f"{name}.xml"
        assert HL.store_frame_xml(p, mono, dep) == 0
        paths.append(p)
    R, T = HL.solvedvo_from_files(paths[0], paths[1], W, H, L, iters, O.K640)
    o = O.align_pair(d["ref_gray"][0], d["ref_depth"][0], d["now_gray"][0], L, iters)
    ang = np.arccos(np.clip((np.trace(R.T @ o["R"]) - 1) / 2, -1, 1))
    assert ang < 1e-5 and np.linalg.norm(T - o["T"]) < 1e-5
