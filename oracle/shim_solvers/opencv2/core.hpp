// Both spellings are included by the reference (model.h uses <opencv2/core/core.hpp>, preemption_empty.h <opencv2/core.hpp>).
#pragma once
#include "core/core.hpp"
