"""Generates tests/golden/opencv_stages.npz: golden vectors for the OpenCV-defined stages of the hot path, produced by
cv2 (the only executable ground truth for those stages in this environment -- the reference pins no OpenCV version and
cannot be built here).  Run from the repo root:  python tests/golden/make_golden.py

Call sites restated: cv::Canny(img,150,100,3,true) / distanceTransform(L2,PRECISE) / normalize(0,255,MINMAX) /
filter2D (src/SolveDVO.cpp:1767-1774, 1077-1093), resize NEAREST / AREA and cvtColor BGR2GRAY
(src/camTopic2PublisherPyD.cpp:344-347, src/EPoseEstimator.cpp:251-253).
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402  (synthetic renderer only)


def stages(img):
    edge = cv2.Canny(img, 150, 100, apertureSize=3, L2gradient=True)
    dt = cv2.distanceTransform(255 - edge, cv2.DIST_L2, cv2.DIST_MASK_PRECISE)
    d2 = np.rint(dt.astype(np.float64) ** 2).astype(np.int32)
    # normalise the correctly-rounded sqrt (cv2's own DT differs by 1 ulp on small images)
    dtn = cv2.normalize(np.sqrt(d2.astype(np.float32)), None, 0.0, 255.0, cv2.NORM_MINMAX)
    kx = np.array([[0, 0, 0], [-.5, 0, .5], [0, 0, 0]], np.float32)
    gx = cv2.filter2D(dtn, cv2.CV_32F, kx)
    gy = cv2.filter2D(dtn, cv2.CV_32F, kx.T.copy())
    return edge, d2, dtn, gx, gy


def main():
    rng = np.random.default_rng(2024)
    out = {}
    imgs = []
    d = O.synth_pair(5, 160, 120, (131.25, 131.25, 79.5, 59.5), bgr=True)
    imgs.append(d["ref_gray"])
    imgs.append(d["now_gray"][::2, ::2].copy())                       # 80x60
    blocks = (rng.integers(0, 2, (9, 12)) * 170 + 30).astype(np.uint8).repeat(8, 0).repeat(8, 1)
    imgs.append(blocks)                                                # 96x72 checker-like
    ramp = np.tile(np.linspace(0, 255, 101).astype(np.uint8), (45, 1))
    ramp[20:30, 40:70] = 255
    imgs.append(ramp)                                                  # 101x45, odd width
    noise = cv2.GaussianBlur(rng.integers(0, 256, (64, 64)).astype(np.uint8), (5, 5), 0)
    imgs.append(noise)
    for i, im in enumerate(imgs):
        edge, d2, dtn, gx, gy = stages(im)
        assert edge.any()
        out[f"img{i}"] = im
        out[f"edge{i}"] = edge
        out[f"d2_{i}"] = d2
        out[f"dtn{i}"] = dtn
        out[f"gx{i}"] = gx
        out[f"gy{i}"] = gy
    out["n_images"] = np.array(len(imgs))
    # resize / colour conversion
    bgr = d["ref_bgr"]                                                  # 120x160x3
    depth = d["ref_depth"]
    out["bgr"] = bgr
    out["depth"] = depth
    out["gray_from_bgr"] = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    for k in (1, 2, 3):
        s = 0.5 ** k
        out[f"nearest_bgr_{k}"] = cv2.resize(bgr, None, fx=s, fy=s, interpolation=cv2.INTER_NEAREST)
        out[f"nearest_depth_{k}"] = cv2.resize(depth, None, fx=s, fy=s, interpolation=cv2.INTER_NEAREST)
        out[f"area_bgr_{k}"] = cv2.resize(bgr, None, fx=s, fy=s, interpolation=cv2.INTER_AREA)
        out[f"area_gray_{k}"] = cv2.resize(out["gray_from_bgr"], None, fx=s, fy=s, interpolation=cv2.INTER_AREA)
        out[f"area_depth_{k}"] = cv2.resize(depth, None, fx=s, fy=s, interpolation=cv2.INTER_AREA)
    # cvRound(dim * scale): 720 / 32 = 22.5 -> 22 (half to even)
    big = np.zeros((720, 1280), np.uint8)
    out["dims_1280x720"] = np.array([cv2.resize(big, None, fx=0.5 ** k, fy=0.5 ** k, interpolation=cv2.INTER_NEAREST).shape for k in range(6)])
    odd = np.zeros((75, 100), np.uint8)
    out["dims_100x75"] = np.array([cv2.resize(odd, None, fx=0.5 ** k, fy=0.5 ** k, interpolation=cv2.INTER_NEAREST).shape for k in range(4)])
    # depth metres -> mm u16 as the pyramid publisher does it (src/camTopic2PublisherPyD.cpp:72-80)
    dm = np.array([[0.0, 0.0004, 0.0005, 0.0015, 1.2345, 2.5005, 65.0, 70.0, 3.9995, 0.0025]], np.float32)
    # (1000.0 * depth).convertTo(CV_16U): fp32 product, then saturate_cast<ushort>(cvRound(x)); cv2.multiply with an
    # explicit dtype takes the same fp32-product + saturate_cast path
    mm16 = cv2.multiply(dm, np.full_like(dm, 1000.0), dtype=cv2.CV_16U)
    mm16[mm16 == 0] = 1
    out["depth_m"] = dm
    out["depth_mm"] = mm16
    np.savez_compressed(os.path.join(HERE, "opencv_stages.npz"), **out)
    print("wrote", os.path.join(HERE, "opencv_stages.npz"), "cv2", cv2.__version__)
    frame_dump()
    undistort_golden()


UND_K = (262.5, 262.5, 159.5, 119.5)
UND_D = (0.2312, -0.7849, -0.0033, -0.0001, 0.9172)        # k1 k2 p1 p2 k3, a Kinect-class lens


def undistort_golden():
    """cv2.undistort of the publisher's two inputs (BGR u8 and 16-bit depth, src/camTopic2PublisherPyD.cpp:86-117) and of a
    gray image, 320x240, so that the oracle's restatement is pinned without cv2 at test time."""
    rng = np.random.default_rng(77)
    d = O.synth_pair(11, 320, 240, UND_K, bgr=True)
    Km = np.array([[UND_K[0], 0, UND_K[2]], [0, UND_K[1], UND_K[3]], [0, 0, 1]]); Dm = np.array(UND_D)
    out = {"K": np.array(UND_K), "D": np.array(UND_D)}
    for name, img in (("bgr", d["ref_bgr"]), ("depth", d["ref_depth"]), ("gray", d["ref_gray"]),
                      ("noise_u8", rng.integers(0, 256, (240, 320), dtype=np.uint8)), ("noise_u16", rng.integers(0, 65536, (240, 320), dtype=np.uint16))):
        out[name + "_in"] = img
        out[name + "_out"] = cv2.undistort(img, Km, Dm)
    np.savez_compressed(os.path.join(HERE, "undistort_320x240.npz"), **out)
    print("wrote undistort_320x240.npz")


def frame_dump():
    """A frame dump exactly as camTopic2PublisherPyD writes it (src/camTopic2PublisherPyD.cpp:325-355): cv::FileStorage XML
    with mono_i / depth_i, level i = cv::resize(full, 0.5^(i+1), INTER_NEAREST); the test parses it without cv2."""
    d = O.synth_pair(9, 128, 96, (105.0, 105.0, 63.5, 47.5), bgr=True)
    bgr, depth = d["ref_bgr"], d["ref_depth"]
    fs = cv2.FileStorage(os.path.join(HERE, "framemono_0000.xml"), cv2.FILE_STORAGE_WRITE)
    want = {}
    scale = 1.0
    for i in range(4):
        scale *= 0.5
        frame = cv2.resize(bgr, None, fx=scale, fy=scale, interpolation=cv2.INTER_NEAREST)
        dframe = cv2.resize(depth, None, fx=scale, fy=scale, interpolation=cv2.INTER_NEAREST)
        mono = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)
        fs.write(f"mono_{i}", mono); fs.write(f"depth_{i}", dframe)
        want[f"mono_{i}"] = mono; want[f"depth_{i}"] = dframe
    fs.release()
    np.savez_compressed(os.path.join(HERE, "framemono_0000.npz"), **want)
    print("wrote framemono_0000.xml / .npz")


if __name__ == "__main__":
    main()
