// PyramidalStorage.h -- per-level cache of the reference frame (reference include/PyramidalStorage.h:37-78,
// src/PyramidalStorage.cpp:11-127).  Same method names and the same eleven members per level.  The reference deep-copies
// 11 host arrays per level in and out (the level-0 Jacobian alone is 14.7 MB at 640x480); here the pyramid lives on the
// device inside a dvo_photo_ctx and addLevel() only records which device level a storage index refers to -- the `level`
// argument is ignored exactly as in the reference (:37-60), entries are addressed in push order.  getLevel()
// materialises the requested level into the caller's host arrays.
#pragma once
#include <vector>

#include "dvo_b200.h"
#include "dvo_types.h"

class PyramidalStorageStruct {
public:
    PyramidalStorageStruct();
    virtual ~PyramidalStorageStruct();
    void bind(dvo_photo_ctx* ctx, int slot, int compat);      // not in the reference: attaches the device storage
    // the eleven references of the original signature carry no data here (the device already holds them); pass the level
    void addLevel(int level, int device_level);
    void getLevel(int level, std::vector<uint8_t>& im_r_color, std::vector<uint8_t>& im_r, std::vector<uint16_t>& dim_r, dvo::ArrayXXd& X,
                  dvo::ArrayXXd& Y, dvo::ArrayXXd& Z, dvo::MatrixXd& J, dvo::ArrayXXd& grayVals, dvo::ArrayXXd& redVals,
                  dvo::ArrayXXd& greenVals, dvo::ArrayXXd& blueVals);
    void clearPyramid();
    void printSize();
    int size() const { return (int)levels_.size(); }
    int deviceLevelAt(int i) const { return levels_.at(i); }
private:
    dvo_photo_ctx* ctx_; int slot_, compat_, width_, height_;
    std::vector<int> levels_;
    friend class EPoseEstimator;
};
