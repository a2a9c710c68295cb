// PyramidalStorage.h -- per-level cache of the reference frame (reference include/PyramidalStorage.h:37-78,
// src/PyramidalStorage.cpp:11-127).  Same method names, same eleven members per level, same semantics: addLevel()
// push_back's (the `level` argument is only asserted, :46), getLevel() returns entry .at(level), clearPyramid(), printSize().
//
// The reference deep-copies eleven host arrays per level in and out (the level-0 Jacobian alone is 14.7 MB at 640x480).
// Here every entry lives on the device:
//   * addLevel(level, im_r_color, im_r, dim_r, X, Y, Z, J, grayVals, redVals, greenVals, blueVals) -- the reference's signature
//     (include/PyramidalStorage.h:42-48) -- uploads the caller's eleven arrays into device blobs owned by this object.  When
//     the storage is bound to an estimator context and the images have that context's level size, im_r / dim_r (and im_r_color
//     at level 0) are also written into the context's pyramid, so EPoseEstimator::estimate consumes the caller's images.
//     The derived members are pure functions of (im_r, dim_r, K, level); the estimator's kernels re-evaluate them from 5 B/px
//     instead of reading 48-byte Jacobian rows back, so what getLevel() returns for them is exactly what was pushed, while the
//     estimator uses its own evaluation of the same functions.
//   * addLevelFromDevice(level, device_level) -- the path EPoseEstimator::setRefFrame takes: nothing is copied, the entry refers
//     to a level the device already holds; getLevel() materialises it on demand.
#pragma once
#include <vector>

#include "dvo_b200.h"
#include "dvo_types.h"

class PyramidalStorageStruct {
public:
    PyramidalStorageStruct();
    virtual ~PyramidalStorageStruct();
    PyramidalStorageStruct(const PyramidalStorageStruct&) = delete;
    PyramidalStorageStruct& operator=(const PyramidalStorageStruct&) = delete;

    void addLevel(int level, const dvo::ImageView& im_r_color, const dvo::ImageView& im_r, const dvo::ImageView& dim_r, dvo::ArrayXXd& X,
                  dvo::ArrayXXd& Y, dvo::ArrayXXd& Z, dvo::MatrixXd& J, dvo::ArrayXXd& grayVals, dvo::ArrayXXd& redVals,
                  dvo::ArrayXXd& greenVals, dvo::ArrayXXd& blueVals);
    void getLevel(int level, std::vector<uint8_t>& im_r_color, std::vector<uint8_t>& im_r, std::vector<uint16_t>& dim_r, dvo::ArrayXXd& X,
                  dvo::ArrayXXd& Y, dvo::ArrayXXd& Z, dvo::MatrixXd& J, dvo::ArrayXXd& grayVals, dvo::ArrayXXd& redVals,
                  dvo::ArrayXXd& greenVals, dvo::ArrayXXd& blueVals);
    void clearPyramid();
    void printSize();
#if defined(DVO_HAVE_OPENCV) && defined(DVO_HAVE_EIGEN)
    // the reference's exact signatures (include/PyramidalStorage.h:42-57)
    void addLevel(int level, cv::Mat& im_r_color, cv::Mat& im_r, cv::Mat& dim_r, Eigen::ArrayXXd& X, Eigen::ArrayXXd& Y, Eigen::ArrayXXd& Z,
                  Eigen::MatrixXd& J, Eigen::ArrayXXd& grayVals, Eigen::ArrayXXd& redVals, Eigen::ArrayXXd& greenVals, Eigen::ArrayXXd& blueVals) {
        dvo::ArrayXXd x, y, z, j, gv, rv, gr, bv;
        dvo::fromEigen(X, x); dvo::fromEigen(Y, y); dvo::fromEigen(Z, z); dvo::fromEigen(J, j);
        dvo::fromEigen(grayVals, gv); dvo::fromEigen(redVals, rv); dvo::fromEigen(greenVals, gr); dvo::fromEigen(blueVals, bv);
        addLevel(level, dvo::ImageView(im_r_color), dvo::ImageView(im_r), dvo::ImageView(dim_r), x, y, z, j, gv, rv, gr, bv);
    }
    void getLevel(int level, cv::Mat& im_r_color, cv::Mat& im_r, cv::Mat& dim_r, Eigen::ArrayXXd& X, Eigen::ArrayXXd& Y, Eigen::ArrayXXd& Z,
                  Eigen::MatrixXd& J, Eigen::ArrayXXd& grayVals, Eigen::ArrayXXd& redVals, Eigen::ArrayXXd& greenVals, Eigen::ArrayXXd& blueVals) {
        std::vector<uint8_t> c, g; std::vector<uint16_t> d; dvo::ArrayXXd x, y, z, j, gv, rv, gr, bv;
        getLevel(level, c, g, d, x, y, z, j, gv, rv, gr, bv);
        const int rows = rowsAt(level), cols = colsAt(level);
        if (!c.empty()) dvo::toMat(c.data(), rows, cols, CV_8UC3, c.size(), im_r_color);
        dvo::toMat(g.data(), rows, cols, CV_8UC1, g.size(), im_r); dvo::toMat(d.data(), rows, cols, CV_16UC1, d.size() * 2, dim_r);
        dvo::toEigen(x, X); dvo::toEigen(y, Y); dvo::toEigen(z, Z); dvo::toEigen(j, J);
        dvo::toEigen(gv, grayVals); dvo::toEigen(rv, redVals); dvo::toEigen(gr, greenVals); dvo::toEigen(bv, blueVals);
    }
#endif

    // ---- not in the reference: attaching the device storage
    void bind(dvo_photo_ctx* ctx, int slot, int compat, int device = 0);
    void addLevelFromDevice(int level, int device_level);
    int size() const { return (int)entries_.size(); }
    int deviceLevelAt(int i) const { return entries_.at(i).device_level; }
    bool isUserLevel(int i) const { return entries_.at(i).user; }
    int rowsAt(int i) const { return entries_.at(i).rows; }
    int colsAt(int i) const { return entries_.at(i).cols; }
private:
    struct Entry {
        bool user = false;             // true: eleven caller-provided arrays in device blobs; false: a level the context computes
        int device_level = -1;         // context level this entry corresponds to (user entries: the level whose size matches, or -1)
        int rows = 0, cols = 0;
        dvo_blob* blob[11] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    };
    dvo_photo_ctx* ctx_; int slot_, compat_, device_, width_, height_;
    std::vector<Entry> entries_;
    void release(Entry& e);
    friend class EPoseEstimator;
};
