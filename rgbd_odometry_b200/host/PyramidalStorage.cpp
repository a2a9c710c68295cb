#include "PyramidalStorage.h"

#include <cassert>
#include <cstdio>
#include <cstdlib>

namespace {
// order of the eleven members (include/PyramidalStorage.h:63-77)
enum { M_COLOR = 0, M_GRAY, M_DEPTH, M_X, M_Y, M_Z, M_J, M_GRAYVALS, M_REDVALS, M_GREENVALS, M_BLUEVALS };
void must(int rc, const char* what) {
    if (rc != DVO_OK) { std::fprintf(stderr, "PyramidalStorageStruct::%s: %s\n", what, dvo_last_error()); assert(rc == DVO_OK); std::abort(); }   // no host fallback
}
}  // namespace

PyramidalStorageStruct::PyramidalStorageStruct() : ctx_(nullptr), slot_(0), compat_(1), device_(0), width_(0), height_(0) { entries_.reserve(5); }   // :11-28
PyramidalStorageStruct::~PyramidalStorageStruct() { clearPyramid(); }
void PyramidalStorageStruct::bind(dvo_photo_ctx* ctx, int slot, int compat, int device) { ctx_ = ctx; slot_ = slot; compat_ = compat; device_ = device; }
void PyramidalStorageStruct::release(Entry& e) { for (int i = 0; i < 11; ++i) { if (e.blob[i]) dvo_blob_destroy(e.blob[i]); e.blob[i] = nullptr; } }

void PyramidalStorageStruct::addLevelFromDevice(int level, int device_level) {
    assert(level >= 0 && level <= 4);                                                      // :46
    Entry e; e.user = false; e.device_level = device_level; e.rows = height_ >> device_level; e.cols = width_ >> device_level;
    entries_.push_back(e);                                                                 // push_back, `level` ignored (:48-60)
}

// uses push_back, "effective level is ignored" (:37): entries are addressed in push order
void PyramidalStorageStruct::addLevel(int level, const dvo::ImageView& im_r_color, const dvo::ImageView& im_r, const dvo::ImageView& dim_r,
                                      dvo::ArrayXXd& X, dvo::ArrayXXd& Y, dvo::ArrayXXd& Z, dvo::MatrixXd& J, dvo::ArrayXXd& grayVals,
                                      dvo::ArrayXXd& redVals, dvo::ArrayXXd& greenVals, dvo::ArrayXXd& blueVals) {
    assert(level >= 0 && level <= 4);                                                      // :46
    assert(im_r.data && dim_r.data && im_r.rows == dim_r.rows && im_r.cols == dim_r.cols);
    Entry e; e.user = true; e.rows = im_r.rows; e.cols = im_r.cols;
    const void* src[11] = {im_r_color.data, im_r.data, dim_r.data, X.data.data(), Y.data.data(), Z.data.data(), J.data.data(),
                           grayVals.data.data(), redVals.data.data(), greenVals.data.data(), blueVals.data.data()};
    const size_t bytes[11] = {im_r_color.data ? im_r_color.bytes() : 0, im_r.bytes(), dim_r.bytes(), X.data.size() * 8, Y.data.size() * 8, Z.data.size() * 8,
                              J.data.size() * 8, grayVals.data.size() * 8, redVals.data.size() * 8, greenVals.data.size() * 8, blueVals.data.size() * 8};
    for (int i = 0; i < 11; ++i) {
        if (bytes[i] == 0) continue;                                                       // an empty member stays empty (getLevel returns it empty)
        must(dvo_blob_create(bytes[i], device_, &e.blob[i]), "addLevel");
        must(dvo_blob_upload(e.blob[i], src[i], bytes[i]), "addLevel");
    }
    // bound to an estimator: the caller's images replace the context's level of the same size, A = J^T J is refreshed
    if (ctx_) {
        for (int dl = 0; dl < 5; ++dl)
            if ((height_ >> dl) == e.rows && (width_ >> dl) == e.cols && (height_ >> dl) > 0) { e.device_level = dl; break; }
        if (e.device_level >= 0) {
            must(dvo_photo_put_level(ctx_, slot_, DVO_FRAME_REF, e.device_level, DVO_PHOTO_GRAY, im_r.data, im_r.bytes()), "addLevel");
            must(dvo_photo_put_level(ctx_, slot_, DVO_FRAME_REF, e.device_level, DVO_PHOTO_DEPTH, dim_r.data, dim_r.bytes()), "addLevel");
            if (e.device_level == 0 && im_r_color.data) must(dvo_photo_put_level(ctx_, slot_, DVO_FRAME_REF, 0, DVO_PHOTO_BGR, im_r_color.data, im_r_color.bytes()), "addLevel");
            must(dvo_photo_prepare_ref(ctx_, slot_, 1, compat_), "addLevel");
        }
    }
    entries_.push_back(e);
}

void PyramidalStorageStruct::clearPyramid() { for (auto& e : entries_) release(e); entries_.clear(); }                                     // :105-123
void PyramidalStorageStruct::printSize() {                                                                                             // :126-127
    int user = 0; for (const auto& e : entries_) user += e.user ? 1 : 0;
    std::printf("PyramidalStorageStruct: %d levels (device resident; %d pushed by the caller)\n", (int)entries_.size(), user);
}

void PyramidalStorageStruct::getLevel(int level, std::vector<uint8_t>& im_r_color, std::vector<uint8_t>& im_r, std::vector<uint16_t>& dim_r,
                                      dvo::ArrayXXd& X, dvo::ArrayXXd& Y, dvo::ArrayXXd& Z, dvo::MatrixXd& J, dvo::ArrayXXd& grayVals,
                                      dvo::ArrayXXd& redVals, dvo::ArrayXXd& greenVals, dvo::ArrayXXd& blueVals) {
    const Entry& e = entries_.at(level);                                                                       // .at(level) (:85-99)
    const int rows = e.rows, cols = e.cols; const size_t P = (size_t)rows * cols;
    dvo::ArrayXXd* arr[7] = {&X, &Y, &Z, &grayVals, &redVals, &greenVals, &blueVals};
    if (e.user) {
        auto fetch = [&](int m, void* dst, size_t cap) { if (e.blob[m]) must(dvo_blob_download(e.blob[m], dst, cap), "getLevel"); };
        im_r_color.resize(e.blob[M_COLOR] ? dvo_blob_bytes(e.blob[M_COLOR]) : 0); fetch(M_COLOR, im_r_color.data(), im_r_color.size());
        im_r.resize(P); fetch(M_GRAY, im_r.data(), P);
        dim_r.resize(P); fetch(M_DEPTH, dim_r.data(), P * 2);
        const int idx[7] = {M_X, M_Y, M_Z, M_GRAYVALS, M_REDVALS, M_GREENVALS, M_BLUEVALS};
        for (int i = 0; i < 7; ++i) {
            const size_t n = e.blob[idx[i]] ? dvo_blob_bytes(e.blob[idx[i]]) / 8 : 0;
            arr[i]->rows = n ? rows : 0; arr[i]->cols = n ? cols : 0; arr[i]->data.resize(n);
            fetch(idx[i], arr[i]->data.data(), n * 8);
        }
        const size_t nj = e.blob[M_J] ? dvo_blob_bytes(e.blob[M_J]) / 8 : 0;
        J.rows = (int)(nj / 6); J.cols = 6; J.data.resize(nj); fetch(M_J, J.data.data(), nj * 8);
        return;
    }
    assert(ctx_ && "storage not bound to a device context");
    const int dl = e.device_level;
    im_r_color.resize(P * 3); im_r.resize(P); dim_r.resize(P);
    int rc = dvo_photo_get_level(ctx_, slot_, DVO_FRAME_REF, dl, DVO_PHOTO_BGR, im_r_color.data(), P * 3, compat_);
    rc |= dvo_photo_get_level(ctx_, slot_, DVO_FRAME_REF, dl, DVO_PHOTO_GRAY, im_r.data(), P, compat_);
    rc |= dvo_photo_get_level(ctx_, slot_, DVO_FRAME_REF, dl, DVO_PHOTO_DEPTH, dim_r.data(), P * 2, compat_);
    const int which[7] = {DVO_PHOTO_X, DVO_PHOTO_Y, DVO_PHOTO_Z, DVO_PHOTO_GRAYVALS, DVO_PHOTO_REDVALS, DVO_PHOTO_GREENVALS, DVO_PHOTO_BLUEVALS};
    for (int i = 0; i < 7; ++i) {
        arr[i]->rows = rows; arr[i]->cols = cols; arr[i]->data.resize(P);
        rc |= dvo_photo_get_level(ctx_, slot_, DVO_FRAME_REF, dl, which[i], arr[i]->data.data(), P * 8, compat_);
    }
    J.rows = (int)P; J.cols = 6; J.data.resize(P * 6);
    rc |= dvo_photo_get_level(ctx_, slot_, DVO_FRAME_REF, dl, DVO_PHOTO_J, J.data.data(), P * 48, compat_);
    must(rc, "getLevel");
}
