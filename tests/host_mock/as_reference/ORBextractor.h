// What a maintainer's O3/include/ORBextractor.h becomes: the adapter class under the reference's header name,
// so that every translation unit that includes "ORBextractor.h" (Frame.cc, Tracking.cc, oracle/cvshim/ref_glue.cpp)
// compiles unchanged against libdvmslam_b200.so.
#pragma once
#include "orb_extractor_adapter.h"
