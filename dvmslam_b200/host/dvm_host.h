// dvm_host.h -- what every C++ adapter over libdvmslam_b200.so shares: status -> exception, device and
// image-size selection, RAII owners of the C handles.  The adapters give back the reference's own class
// signatures (ORB_SLAM3::ORBextractor / ORBmatcher / Optimizer, O3/ = src/slam_system/orb_slam3/) so that
// Tracking.cc, LocalMapping.cc and the ROS2 wrapper compile unchanged; see INTEGRATION.md.
#pragma once
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "dvmslam_b200.h"

namespace dvm_host {

// The reference's operators do not return errors (they assert or crash); a failed GPU call therefore
// surfaces as an exception at the call site instead of being swallowed.  There is no CPU fallback.
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) { }
};

inline void check(int rc, const char* where)
{
    if (rc != DVM_OK) throw Error(rc, std::string(where) + ": " + dvm_last_error());
}

inline int env_int(const char* name, int fallback)
{
    const char* v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : fallback;
}

// one agent = one process = one GPU: the ordinal comes from DVM_DEVICE, else LOCAL_RANK (torchrun), else 0
inline int device_from_env() { return env_int("DVM_DEVICE", env_int("LOCAL_RANK", 0)); }
// device buffers are sized once per extractor; larger images are rejected by dvm_orb_extract
inline int max_width() { return env_int("DVM_MAX_WIDTH", 1920); }
inline int max_height() { return env_int("DVM_MAX_HEIGHT", 1200); }

struct OrbHandle {
    dvm_orb* h = nullptr;
    OrbHandle() = default;
    OrbHandle(const OrbHandle&) = delete;
    OrbHandle& operator=(const OrbHandle&) = delete;
    ~OrbHandle() { dvm_orb_destroy(h); }
};

struct FrameHandle {
    dvm_frame* h = nullptr;
    FrameHandle() = default;
    FrameHandle(const FrameHandle&) = delete;
    FrameHandle& operator=(const FrameHandle&) = delete;
    ~FrameHandle() { dvm_frame_destroy(h); }
};

struct LbaHandle {
    dvm_lba* h = nullptr;
    LbaHandle() = default;
    LbaHandle(const LbaHandle&) = delete;
    LbaHandle& operator=(const LbaHandle&) = delete;
    ~LbaHandle() { dvm_lba_destroy(h); }
};

struct Sim3Handle {
    dvm_sim3* h = nullptr;
    Sim3Handle() = default;
    Sim3Handle(const Sim3Handle&) = delete;
    Sim3Handle& operator=(const Sim3Handle&) = delete;
    ~Sim3Handle() { dvm_sim3_destroy(h); }
};

} // namespace dvm_host
