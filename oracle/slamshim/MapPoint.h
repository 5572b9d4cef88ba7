/* TEST INFRASTRUCTURE: stands in for O3/include/MapPoint.h when the reference's ORBmatcher.{h,cc} are compiled through oracle/_ref/tree (see slam_types.h). */
#pragma once
#include "slam_types.h"
