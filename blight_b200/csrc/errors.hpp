// errors.hpp — internal aliases of the status codes in include/blight_b200.h.
#pragma once
#include "../../include/blight_b200.h"

namespace blight {
constexpr int BL_OK = BLIGHT_OK;
constexpr int BL_ERR_INVALID_ARG = BLIGHT_ERR_INVALID_ARG;
constexpr int BL_ERR_IO = BLIGHT_ERR_IO;
constexpr int BL_ERR_INVALID_BASE = BLIGHT_ERR_INVALID_BASE;
constexpr int BL_ERR_CUDA = BLIGHT_ERR_CUDA;
constexpr int BL_ERR_NO_DEVICE = BLIGHT_ERR_NO_DEVICE;
constexpr int BL_ERR_FORMAT = BLIGHT_ERR_FORMAT;
constexpr int BL_ERR_NOMEM = BLIGHT_ERR_NOMEM;
}  // namespace blight
