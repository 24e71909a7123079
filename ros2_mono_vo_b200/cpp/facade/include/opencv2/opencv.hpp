// opencv2/opencv.hpp -- umbrella header of the OpenCV-API facade (see opencv2/core.hpp).  The real umbrella header
// drags in most of the standard library; the reference relies on that (std::map, std::unordered_map and assert are
// used without their own includes in map.hpp / keyframe.hpp / keyframe.cpp), so the same headers come along here.
#pragma once
#include <algorithm>
#include <cassert>
#include <iostream>
#include <map>
#include <unordered_map>
#include <utility>

#include "opencv2/core.hpp"
