// FrameIO.cpp -- see FrameIO.h.  A dependency-free reader / writer for the subset of cv::FileStorage XML the
// reference uses (opencv-matrix nodes), the pose text lines and the TUM ground-truth file.
#include "FrameIO.h"

#include "dvo_b200.h"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

namespace dvo {

namespace {

bool slurp(const char* fileName, std::string& s) {
    std::ifstream f(fileName, std::ios::binary);
    if (!f.is_open()) return false;
    std::stringstream ss; ss << f.rdbuf(); s = ss.str();
    return true;
}

// text of <tag>...</tag> searched in [from, to); npos-safe
bool inner(const std::string& s, const std::string& tag, size_t from, size_t to, std::string& out) {
    const size_t a = s.find("<" + tag + ">", from);
    if (a == std::string::npos || a >= to) return false;
    const size_t b = s.find("</" + tag + ">", a);
    if (b == std::string::npos || b > to) return false;
    out = s.substr(a + tag.size() + 2, b - a - tag.size() - 2);
    return true;
}

// cv::FileStorage dt strings: [count]<u|c|w|s|i|f|d>, quoted when a count is present ("3u")
bool parseDt(std::string dt, int& channels, int& elem) {
    std::string t;
    for (char ch : dt) if (ch != '"' && ch != ' ' && ch != '\n' && ch != '\t' && ch != '\r') t.push_back(ch);
    if (t.empty()) return false;
    channels = 1;
    size_t k = 0;
    if (t[0] >= '0' && t[0] <= '9') { channels = std::atoi(t.c_str()); while (k < t.size() && t[k] >= '0' && t[k] <= '9') ++k; }
    if (k >= t.size() || channels < 1) return false;
    switch (t[k]) {
        case 'u': elem = ELEM_U8; break;
        case 'w': elem = ELEM_U16; break;
        case 's': elem = ELEM_S16; break;
        case 'i': elem = ELEM_S32; break;
        case 'f': elem = ELEM_F32; break;
        case 'd': elem = ELEM_F64; break;
        default: return false;
    }
    return true;
}

const char* dtChar(int elem) {
    switch (elem) { case ELEM_U8: return "u"; case ELEM_U16: return "w"; case ELEM_S16: return "s"; case ELEM_S32: return "i"; case ELEM_F32: return "f"; default: return "d"; }
}

template <typename T> void rotFromQuat(T w, T x, T y, T z, Mat3<T>& R) {      // Eigen::Quaternion::toRotationMatrix (src/loadGTPath.cpp:133-134)
    const T tx = T(2) * x, ty = T(2) * y, tz = T(2) * z;
    const T twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R(0, 0) = T(1) - (tyy + tzz); R(0, 1) = txy - twz; R(0, 2) = txz + twy;
    R(1, 0) = txy + twz; R(1, 1) = T(1) - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy; R(2, 1) = tyz + twx; R(2, 2) = T(1) - (txx + tyy);
}

void quatFromRot(const Matrix3f& R, float* q) {                              // Eigen::Quaternionf(rot) (src/loadGTPath.cpp:37)
    const float tr = R(0, 0) + R(1, 1) + R(2, 2);
    if (tr > 0) {
        float s = std::sqrt(tr + 1.0f); q[3] = 0.5f * s; s = 0.5f / s;
        q[0] = (R(2, 1) - R(1, 2)) * s; q[1] = (R(0, 2) - R(2, 0)) * s; q[2] = (R(1, 0) - R(0, 1)) * s;
    } else {
        int i = 0; if (R(1, 1) > R(0, 0)) i = 1; if (R(2, 2) > R(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        float s = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0f); q[i] = 0.5f * s; s = 0.5f / s;
        q[3] = (R(k, j) - R(j, k)) * s; q[j] = (R(j, i) + R(i, j)) * s; q[k] = (R(k, i) + R(i, k)) * s;
    }
}

void ltrim(std::string& line) {                                              // src/loadGTPath.cpp:115, :127
    size_t k = 0;
    while (k < line.size() && std::isspace((unsigned char)line[k])) ++k;
    line.erase(0, k);
}

}  // namespace

bool readXmlMatrix(const std::string& xml, const std::string& name, Image& out) {
    const size_t a = xml.find("<" + name + " ");
    if (a == std::string::npos) return false;
    const size_t b = xml.find("</" + name + ">", a);
    if (b == std::string::npos) return false;
    const size_t gt = xml.find('>', a);
    if (gt == std::string::npos || gt > b || xml.substr(a, gt - a).find("opencv-matrix") == std::string::npos) return false;
    std::string rows, cols, dt, data;
    if (!inner(xml, "rows", a, b, rows) || !inner(xml, "cols", a, b, cols) || !inner(xml, "dt", a, b, dt) || !inner(xml, "data", a, b, data)) return false;
    Image m;
    m.rows = std::atoi(rows.c_str()); m.cols = std::atoi(cols.c_str());
    if (m.rows < 0 || m.cols < 0 || !parseDt(dt, m.channels, m.elem)) return false;
    const size_t n = (size_t)m.rows * m.cols * m.channels;
    m.data.resize(n * m.elemSize());
    const char* p = data.c_str();
    char* end = nullptr;
    for (size_t i = 0; i < n; ++i) {
        if (m.elem == ELEM_F32 || m.elem == ELEM_F64) {
            const double v = std::strtod(p, &end);
            if (end == p) return false;
            if (m.elem == ELEM_F32) reinterpret_cast<float*>(m.data.data())[i] = (float)v; else reinterpret_cast<double*>(m.data.data())[i] = v;
        } else {
            const long v = std::strtol(p, &end, 10);
            if (end == p) return false;
            switch (m.elem) {
                case ELEM_U8: m.data[i] = (uint8_t)v; break;
                case ELEM_U16: reinterpret_cast<uint16_t*>(m.data.data())[i] = (uint16_t)v; break;
                case ELEM_S16: reinterpret_cast<int16_t*>(m.data.data())[i] = (int16_t)v; break;
                default: reinterpret_cast<int32_t*>(m.data.data())[i] = (int32_t)v; break;
            }
        }
        p = end;
    }
    out = std::move(m);
    return true;
}

void writeXmlMatrix(std::ostream& os, const std::string& name, const Image& m) {
    os << "<" << name << " type_id=\"opencv-matrix\">\n  <rows>" << m.rows << "</rows>\n  <cols>" << m.cols << "</cols>\n  <dt>";
    if (m.channels > 1) os << "\"" << m.channels << dtChar(m.elem) << "\""; else os << dtChar(m.elem);
    os << "</dt>\n  <data>\n    ";
    const size_t n = (size_t)m.rows * m.cols * m.channels;
    char buf[40];
    size_t col = 4;
    for (size_t i = 0; i < n; ++i) {
        int len;
        switch (m.elem) {
            case ELEM_U8: len = std::snprintf(buf, sizeof buf, "%u", (unsigned)m.data[i]); break;
            case ELEM_U16: len = std::snprintf(buf, sizeof buf, "%u", (unsigned)reinterpret_cast<const uint16_t*>(m.data.data())[i]); break;
            case ELEM_S16: len = std::snprintf(buf, sizeof buf, "%d", (int)reinterpret_cast<const int16_t*>(m.data.data())[i]); break;
            case ELEM_S32: len = std::snprintf(buf, sizeof buf, "%d", (int)reinterpret_cast<const int32_t*>(m.data.data())[i]); break;
            case ELEM_F32: len = std::snprintf(buf, sizeof buf, "%.9g", (double)reinterpret_cast<const float*>(m.data.data())[i]); break;      // round-trips fp32
            default: len = std::snprintf(buf, sizeof buf, "%.17g", reinterpret_cast<const double*>(m.data.data())[i]); break;
        }
        if (i > 0) { if (col + 1 + (size_t)len > 72) { os << "\n    "; col = 4; } else { os << ' '; ++col; } }
        os << buf; col += (size_t)len;
    }
    os << "</data></" << name << ">\n";
}

bool loadFrameXml(const char* xmlFileName, RGBDFramePyd& out, int levels) {
    std::string xml;
    if (!slurp(xmlFileName, xml)) { std::fprintf(stderr, "[loadFrameXml] Cannot Open File %s\n", xmlFileName); return false; }   // :158-162
    out.framergb.clear(); out.framemono.clear(); out.dframe.clear();                                                              // :164-165
    for (int i = 0; i < levels; ++i) {                                                                                            // :166-186
        Image mono, depth;
        char nm[100], nd[100];
        std::snprintf(nm, sizeof nm, "mono_%d", i); std::snprintf(nd, sizeof nd, "depth_%d", i);
        if (!readXmlMatrix(xml, nm, mono) || !readXmlMatrix(xml, nd, depth)) { std::fprintf(stderr, "[loadFrameXml] %s: node %s / %s missing\n", xmlFileName, nm, nd); return false; }
        out.framemono.push_back(std::move(mono)); out.dframe.push_back(std::move(depth));
    }
    return true;
}

bool storeFrameXml(const char* xmlFileName, const RGBDFramePyd& in) {
    if (in.framemono.size() != in.dframe.size()) return false;
    std::ofstream f(xmlFileName, std::ios::binary);
    if (!f.is_open()) return false;
    f << "<?xml version=\"1.0\"?>\n<opencv_storage>\n";
    for (size_t i = 0; i < in.framemono.size(); ++i) {                           // src/camTopic2PublisherPyD.cpp:352-355
        writeXmlMatrix(f, "mono_" + std::to_string(i), in.framemono[i]);
        writeXmlMatrix(f, "depth_" + std::to_string(i), in.dframe[i]);
    }
    f << "</opencv_storage>\n";
    return f.good();
}

bool undistortFrame(const ImageView& src, Image& dst, const double* K4, const double* D5) {
    dst.rows = src.rows; dst.cols = src.cols; dst.channels = src.channels(); dst.elem = (src.type == U16C1) ? ELEM_U16 : ELEM_U8;
    dst.data.resize(src.bytes());
    const int type = (src.type == U16C1) ? DVO_IMG_U16C1 : (src.type == U8C3 ? DVO_IMG_U8C3 : DVO_IMG_U8C1);
    return dvo_undistort(src.data, dst.data.data(), src.cols, src.rows, type, 1, K4, D5, DVO_MEM_HOST, nullptr) == DVO_OK;
}

void printPose(const Pose& p, std::ostream& stream) {                             // src/SolveDVO.cpp:1346-1350
    stream << p.orientation.x << " " << p.orientation.y << " " << p.orientation.z << " " << p.orientation.w << " "
           << p.position.x << " " << p.position.y << " " << p.position.z << "\n";
    stream.flush();
}

bool readPoseFile(const char* fileName, std::vector<Pose>& out) {
    std::ifstream f(fileName);
    if (!f.is_open()) return false;
    out.clear();
    std::string line;
    while (std::getline(f, line)) {
        ltrim(line);
        if (line.empty() || line[0] == '#') continue;
        std::stringstream ss(line);
        Pose p;
        if (!(ss >> p.orientation.x >> p.orientation.y >> p.orientation.z >> p.orientation.w >> p.position.x >> p.position.y >> p.position.z)) return false;
        out.push_back(p);
    }
    return true;
}

bool loadGTPath(const char* fileName, std::vector<Pose>& out, int skipLines) {
    std::ifstream fin(fileName);
    if (!fin.is_open()) { std::fprintf(stderr, "[loadGTPath] Cannot open file '%s'\n", fileName); return false; }                 // :64-69
    out.clear();
    std::string line;
    if (skipLines >= 0) {                                                         // :110-120 (breaks after consuming line skipLines + 1)
        int tmpCount = 0;
        while (std::getline(fin, line)) {
            ltrim(line);
            if (!line.empty() && line[0] == '#') continue;
            if (++tmpCount > skipLines) break;
        }
    }
    Matrix3f Rf, Rc; Vector3f Tf, Tc;
    int lineCount = 0;
    while (std::getline(fin, line)) {                                             // :123-178
        ltrim(line);
        if (line.empty() || line[0] == '#') continue;                              // (blank lines skipped; the reference would parse zeros)
        float ln[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        std::stringstream ss(line);
        for (int i = 0; i < 8; ++i) { float tmp = 0.f; ss >> tmp; ln[i] = tmp; }   // parseData :18-29 (fp32, also for the time stamp)
        rotFromQuat<float>(ln[7], ln[4], ln[5], ln[6], Rc);                        // Quaternionf(w, x, y, z) :141
        Tc = Vector3f(ln[1], ln[2], ln[3]);
        if (lineCount == 0) { Rf = Rc; Tf = Tc; }                                  // :131-138
        Matrix3f RfT;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) RfT(r, c) = Rf(c, r);
        const Vector3f Tu = RfT * Vector3f(Tc(0) - Tf(0), Tc(1) - Tf(1), Tc(2) - Tf(2));   // :149
        const Matrix3f Ru = RfT * Rc;                                              // :150
        float q[4]; quatFromRot(Ru, q);                                            // matrixToPose :35-46
        Pose p;
        p.position.x = Tu(0); p.position.y = Tu(1); p.position.z = Tu(2);
        p.orientation.x = q[0]; p.orientation.y = q[1]; p.orientation.z = q[2]; p.orientation.w = q[3];
        out.push_back(p);
        ++lineCount;
    }
    return true;
}

}  // namespace dvo
