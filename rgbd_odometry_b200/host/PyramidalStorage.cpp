#include "PyramidalStorage.h"

#include <cassert>
#include <cstdio>

PyramidalStorageStruct::PyramidalStorageStruct() : ctx_(nullptr), slot_(0), compat_(1), width_(0), height_(0) { levels_.reserve(5); }   // :11-28
PyramidalStorageStruct::~PyramidalStorageStruct() {}
void PyramidalStorageStruct::bind(dvo_photo_ctx* ctx, int slot, int compat) { ctx_ = ctx; slot_ = slot; compat_ = compat; }
void PyramidalStorageStruct::addLevel(int /*level*/, int device_level) { levels_.push_back(device_level); }    // push_back, level ignored (:38-65)
void PyramidalStorageStruct::clearPyramid() { levels_.clear(); }                                               // :105-123
void PyramidalStorageStruct::printSize() { std::printf("PyramidalStorageStruct: %d levels (device resident)\n", (int)levels_.size()); }

void PyramidalStorageStruct::getLevel(int level, std::vector<uint8_t>& im_r_color, std::vector<uint8_t>& im_r, std::vector<uint16_t>& dim_r,
                                      dvo::ArrayXXd& X, dvo::ArrayXXd& Y, dvo::ArrayXXd& Z, dvo::MatrixXd& J, dvo::ArrayXXd& grayVals,
                                      dvo::ArrayXXd& redVals, dvo::ArrayXXd& greenVals, dvo::ArrayXXd& blueVals) {
    assert(ctx_ && "storage not bound to a device context");
    const int dl = levels_.at(level);                                                                          // .at(level) (:85-99)
    const int rows = height_ >> dl, cols = width_ >> dl; const size_t P = (size_t)rows * cols;
    im_r_color.resize(P * 3); im_r.resize(P); dim_r.resize(P);
    int rc = dvo_photo_get_level(ctx_, slot_, DVO_FRAME_REF, dl, DVO_PHOTO_BGR, im_r_color.data(), P * 3, compat_);
    rc |= dvo_photo_get_level(ctx_, slot_, DVO_FRAME_REF, dl, DVO_PHOTO_GRAY, im_r.data(), P, compat_);
    rc |= dvo_photo_get_level(ctx_, slot_, DVO_FRAME_REF, dl, DVO_PHOTO_DEPTH, dim_r.data(), P * 2, compat_);
    dvo::ArrayXXd* arr[7] = {&X, &Y, &Z, &grayVals, &redVals, &greenVals, &blueVals};
    const int which[7] = {DVO_PHOTO_X, DVO_PHOTO_Y, DVO_PHOTO_Z, DVO_PHOTO_GRAYVALS, DVO_PHOTO_REDVALS, DVO_PHOTO_GREENVALS, DVO_PHOTO_BLUEVALS};
    for (int i = 0; i < 7; ++i) {
        arr[i]->rows = rows; arr[i]->cols = cols; arr[i]->data.resize(P);
        rc |= dvo_photo_get_level(ctx_, slot_, DVO_FRAME_REF, dl, which[i], arr[i]->data.data(), P * 8, compat_);
    }
    J.rows = (int)P; J.cols = 6; J.data.resize(P * 6);
    rc |= dvo_photo_get_level(ctx_, slot_, DVO_FRAME_REF, dl, DVO_PHOTO_J, J.data.data(), P * 48, compat_);
    assert(rc == 0 && "dvo_photo_get_level failed");
    (void)rc;
}
