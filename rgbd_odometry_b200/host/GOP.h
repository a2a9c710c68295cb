// GOP.h -- "group of pictures": global poses from key-frame-relative poses (reference include/GOP.h:67-95,
// src/GOP.cpp:5-245).  Same class and method names; Eigen / geometry_msgs types replaced by dvo_types.h PODs.
// The per-frame bookkeeping is a handful of 3x3 products on the host; batched composition of many sequences runs on
// the device through dvo_gop_compose (include/dvo_b200.h).
#pragma once
#include <cassert>
#include <cmath>
#include <vector>

#include "dvo_types.h"

template <typename T>
class GOPElement {
public:
    GOPElement() : keyFrame(false), frameId(-1), reason_of_change(-1) {}                      // src/GOP.cpp:5-17
    void setAsOrdinaryFrame(int frameNum, dvo::Mat3<T> wR, dvo::Vec3<T> wT) {                  // :24-35
        keyFrame = false; reason_of_change = -1; frameId = frameNum; world_R = wR; world_T = wT; matrixToPose(wR, wT, world_pose);
    }
    void setAsKeyFrame(int frameNum, int reasonCode, dvo::Mat3<T> wR, dvo::Vec3<T> wT) {       // :44-55
        keyFrame = true; reason_of_change = reasonCode; frameId = frameNum; world_R = wR; world_T = wT; matrixToPose(wR, wT, world_pose);
    }
    const dvo::Mat3<T>& getR() { return world_R; }
    const dvo::Vec3<T>& getT() { return world_T; }
    const dvo::Pose& getPose() { return world_pose; }
    bool isKeyFrame() { return keyFrame; }
    int getReason() { return keyFrame ? reason_of_change : -1; }                               // :82-88
    void updateAsKeyFrame(int reason) { keyFrame = true; reason_of_change = reason; }          // :91-95
private:
    bool keyFrame; int frameId; int reason_of_change;
    dvo::Mat3<T> world_R; dvo::Vec3<T> world_T; dvo::Pose world_pose;
    // Eigen::Quaternion<T>(rot) (src/GOP.cpp:103-114)
    void matrixToPose(dvo::Mat3<T> R, dvo::Vec3<T> t, dvo::Pose& p) {
        T q[4];
        const T tr = R(0, 0) + R(1, 1) + R(2, 2);
        if (tr > 0) {
            T s = std::sqrt(tr + T(1)); q[3] = T(0.5) * s; s = T(0.5) / s;
            q[0] = (R(2, 1) - R(1, 2)) * s; q[1] = (R(0, 2) - R(2, 0)) * s; q[2] = (R(1, 0) - R(0, 1)) * s;
        } else {
            int i = 0; if (R(1, 1) > R(0, 0)) i = 1; if (R(2, 2) > R(i, i)) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            T s = std::sqrt(R(i, i) - R(j, j) - R(k, k) + T(1)); q[i] = T(0.5) * s; s = T(0.5) / s;
            q[3] = (R(k, j) - R(j, k)) * s; q[j] = (R(j, i) + R(i, j)) * s; q[k] = (R(k, i) + R(i, k)) * s;
        }
        p.position.x = t(0); p.position.y = t(1); p.position.z = t(2);
        p.orientation.x = q[0]; p.orientation.y = q[1]; p.orientation.z = q[2]; p.orientation.w = q[3];
    }
};

template <typename T>
class GOP {
public:
    GOP() { gopVector.reserve(100000); gopVector.clear(); }                                   // :122-129
    void pushAsOrdinaryFrame(int frameNum, dvo::Mat3<T> cR, dvo::Vec3<T> cT) {                 // :138-154
        dvo::Vec3<T> global_T = lastKeyFr_T + lastKeyFr_R * cT;
        dvo::Mat3<T> global_R = lastKeyFr_R * cR;
        GOPElement<T> ele; ele.setAsOrdinaryFrame(frameNum, global_R, global_T); gopVector.push_back(ele);
    }
    void pushAsKeyFrame(int frameNum, int reason, dvo::Mat3<T> cR, dvo::Vec3<T> cT) {          // :164-186
        dvo::Vec3<T> global_T = lastKeyFr_T + lastKeyFr_R * cT;
        dvo::Mat3<T> global_R = lastKeyFr_R * cR;
        GOPElement<T> ele; ele.setAsKeyFrame(frameNum, reason, global_R, global_T); gopVector.push_back(ele);
        lastKeyFr_R = global_R; lastKeyFr_T = global_T;
    }
    void updateMostRecentToKeyFrame(int reason) {                                             // :189-196
        const int i = (int)gopVector.size() - 1;
        lastKeyFr_R = gopVector[i].getR(); lastKeyFr_T = gopVector[i].getT(); gopVector[i].updateAsKeyFrame(reason);
    }
    int size() { return (int)gopVector.size(); }
    const dvo::Pose& getGlobalPoseAt(int i) { assert(i < (int)gopVector.size()); return gopVector[i].getPose(); }
    const dvo::Mat3<T>& getGlobalRAt(int i) { assert(i < (int)gopVector.size()); return gopVector[i].getR(); }
    const dvo::Vec3<T>& getGlobalTAt(int i) { assert(i < (int)gopVector.size()); return gopVector[i].getT(); }
    bool isKeyFrameAt(int i) { assert(i < (int)gopVector.size()); return gopVector[i].isKeyFrame(); }
    int getReasonAt(int i) { assert(i < (int)gopVector.size()); return gopVector[i].getReason(); }
private:
    std::vector<GOPElement<T> > gopVector;
    dvo::Mat3<T> lastKeyFr_R; dvo::Vec3<T> lastKeyFr_T;
};
