// FrameIO.h -- the on-disk / wire formats either side of the alignment path, without OpenCV or ROS.
//   * frame dumps: OpenCV FileStorage XML with nodes mono_<i> / depth_<i>, one pair per pyramid level, written by
//     camTopic2PublisherPyD (src/camTopic2PublisherPyD.cpp:315-365) and replayed by SolveDVO::loadFromFile
//     (src/SolveDVO.cpp:154-190);
//   * RGBDFramePyd (msg/RGBDFramePyd.msg:1-3): the three per-level image arrays the publisher sends;
//   * pose text files, one "qx qy qz qw tx ty tz" line per frame (SolveDVO::printPose, src/SolveDVO.cpp:1340-1351);
//   * TUM RGB-D ground-truth trajectories "timestamp tx ty tz qx qy qz qw" re-expressed in the first frame read
//     (src/loadGTPath.cpp:18-29, :50-184).
#pragma once
#include <iosfwd>
#include <string>
#include <vector>

#include "dvo_types.h"

namespace dvo {

enum ElemType { ELEM_U8 = 0, ELEM_U16 = 1, ELEM_S16 = 2, ELEM_S32 = 3, ELEM_F32 = 4, ELEM_F64 = 5 };

// an owned, tightly packed row-major matrix (the cv::Mat a FileStorage node holds)
struct Image {
    int rows = 0, cols = 0, channels = 1, elem = ELEM_U8;
    std::vector<uint8_t> data;
    size_t elemSize() const { return elem == ELEM_U8 ? 1 : (elem == ELEM_U16 || elem == ELEM_S16) ? 2 : (elem == ELEM_F64 ? 8 : 4); }
    // view for the class API; only the types the alignment path ingests have an ImageType
    ImageView view() const { return ImageView(data.data(), rows, cols, elem == ELEM_U16 ? U16C1 : (channels == 3 ? U8C3 : U8C1)); }
};

// msg/RGBDFramePyd.msg: sensor_msgs/Image[] framergb, framemono, dframe (index = pyramid level)
struct RGBDFramePyd { std::vector<Image> framergb, framemono, dframe; };

// One named opencv-matrix node of a FileStorage XML document; false when the node is absent or malformed.
bool readXmlMatrix(const std::string& xml, const std::string& name, Image& out);
// Appends a node in the layout cv::FileStorage writes (and reads back).
void writeXmlMatrix(std::ostream& os, const std::string& name, const Image& m);

// SolveDVO::loadFromFile (src/SolveDVO.cpp:154-190): mono_0..mono_{levels-1} and depth_0.. into framemono / dframe.
// Returns false (and logs, like the reference) when the file cannot be opened or a node is missing.
bool loadFrameXml(const char* xmlFileName, RGBDFramePyd& out, int levels = 4);
// the writer side (src/camTopic2PublisherPyD.cpp:315-365): fs << "mono_i" << framemono << "depth_i" << dframe per level
bool storeFrameXml(const char* xmlFileName, const RGBDFramePyd& in);

// undistortFrame / undistortDFrame of the publisher (src/camTopic2PublisherPyD.cpp:86-117): cv::undistort(src, dst,
// cameraMatrix, distCoeff) on the GPU (dvo_undistort; bilinear, constant border, bit-exact against OpenCV 4.13).  K = fx, fy,
// cx, cy; D = k1, k2, p1, p2, k3.  Returns false when the device call fails (the reference copies src when it has no camera info).
bool undistortFrame(const ImageView& src, Image& dst, const double* K4, const double* D5);

// SolveDVO::printPose's file line (src/SolveDVO.cpp:1346-1350): default ostream formatting, space separated, '\n'.
void printPose(const Pose& p, std::ostream& stream);
bool readPoseFile(const char* fileName, std::vector<Pose>& out);

// loadGTPath (src/loadGTPath.cpp:50-184): skip '#' lines, drop the first `skipLines` data lines (the reference
// hard-codes 350, :112-120, and consumes one more line while doing so), parse 8 floats per line (:18-29) and express
// every pose in the frame of the first one kept: Tu = Rf^T (Tc - Tf), Ru = Rf^T Rc, in fp32 like the reference.
bool loadGTPath(const char* fileName, std::vector<Pose>& out, int skipLines = 350);

}  // namespace dvo
