// oracle/shim_solvers/opencv2/core/core.hpp -- TEST INFRASTRUCTURE ONLY: the cv::Mat container shim of oracle/shim.
#pragma once
#include "../../../shim/opencv2/core/core.hpp"
