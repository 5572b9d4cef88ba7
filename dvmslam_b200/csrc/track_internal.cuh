// track_internal.cuh -- the dvm_frame object shared by track_capi.cu and track_pipeline.cu.
#pragma once
#include "track_kernels.cuh"

using dvm::FrameDev;
using dvm::MatchScratch;

struct dvm_frame {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev = nullptr;
    int cap = 0;
    FrameDev dev;
    // device storage
    dvm_keypoint* d_kps = nullptr;
    uint8_t* d_desc = nullptr;
    int* d_n = nullptr;
    int* d_cell_start = nullptr;
    int* d_cell_items = nullptr;
    int* d_kxyo = nullptr;
    // matcher / optimiser scratch, grown on demand
    uint8_t* d_in = nullptr;   // uploaded inputs
    size_t in_cap = 0;
    uint8_t* h_in = nullptr;   // pinned staging for uploads
    size_t h_in_cap = 0;
    MatchScratch ms = {};
    int q_cap = 0;
    int* d_cur_mp = nullptr;   // [cap + 8]: cur_mp, then {nmatches}
    double* d_err = nullptr;
    int err_cap = 0;
    uint8_t* h_out = nullptr;  // pinned read-back
    size_t h_out_cap = 0;
    int last_rounds = 0;
    int host_n = 0;
    bool undistort = false;      // k1 != 0: UndistortKeyPoints is not the identity
    dvm::UndistortArgs und;
};


int dvm_frame_ensure_query_cap(dvm_frame* f, int nq);
int dvm_frame_ensure_bytes(dvm_frame* f, size_t in_bytes, size_t out_bytes);
// reads cur_mp[host_n] / nmatches of the last matcher launch back (synchronises the frame's stream)
int dvm_frame_finish_match(dvm_frame* f, int32_t* cur_mp, int* nmatches);
