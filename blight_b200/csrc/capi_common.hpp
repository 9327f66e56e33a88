// capi_common.hpp — definitions shared by the two halves of the C ABI (capi_host.cpp, capi_device.cu).
#pragma once
#include <string>

#include "errors.hpp"
#include "flat_index.hpp"

struct blight_flat { blight::FlatIndex f; };
struct blight_index;

namespace blight {
extern thread_local std::string g_last_error;
int fail(int code, const std::string& msg);
void fill_info(const FlatIndex& f, blight_info* out);
// Keeps MPHF groups [g_begin, g_end): other buckets become empty, arrays are compacted, ids stay global.
int flat_slice(const FlatIndex& f, uint64_t g_begin, uint64_t g_end, FlatIndex& out, std::string* err);
// stream_query.cu: file_query(path) chunk by chunk (reader thread -> parallel record cut -> H2D / kernel overlap)
int stream_file_query(const struct ::blight_index* idx, const char* path, uint64_t* ctr);
void stream_ctx_free(void* ctx);
}  // namespace blight
