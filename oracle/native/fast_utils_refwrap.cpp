// ORACLE - TEST INFRASTRUCTURE ONLY.  extern "C" doorway to the reference's own native grouping code, which is
// compiled from /root/reference where it lies (never copied): nano_demo/fast_utils/parse/find_peaks.cpp and
// assign.cpp are passed to g++ next to this file by oracle/native/build_native.py; output oracle/_ref/ only.
#include "assign.hpp"
#include "find_peaks.hpp"

extern "C" void ref_find_peaks_nchw(int* count, float* val, float* tag, int* ind, const float* input, const float* tmap,
                                    int N, int C, int H, int W, int M, float threshold, int window_size) {
    trt_pose::parse::find_peaks_out_nchw(count, val, tag, ind, input, tmap, N, C, H, W, M, threshold, window_size);
}

// one image; only defined while every per-joint count and the person count stay <= 10 (the reference's stack arrays)
extern "C" void ref_assign(int* num_person, float* ans, const int* cnt, const float* val, const float* tag, const int* ind,
                           const int* joint_order, int C, int M, float threshold) {
    trt_pose::parse::assign_out(num_person, ans, cnt, val, tag, ind, joint_order, C, M, threshold);
}
