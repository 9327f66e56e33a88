// capi_common.hpp — definitions shared by the two halves of the C ABI (capi_host.cpp, capi_device.cu).
#pragma once
#include <functional>
#include <string>
#include <vector>

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
// gpu_builder.cu: construct_index on the GPU; sequences are views [starts[i], starts[i] + lens[i]) of text (host memory)
int build_flat_index_gpu(const char* text, uint64_t text_len, const std::vector<uint64_t>& starts, const std::vector<uint64_t>& lens,
                         const BuildParams& P, int device, FlatIndex& out, std::string* err, double* seconds_device);
// stream_query.cu: file_query(path) chunk by chunk (reader thread -> parallel record cut -> H2D / kernel overlap)
int stream_file_query(const struct ::blight_index* idx, const char* path, uint64_t* ctr);
void stream_ctx_free(void* ctx);
// the same reader and record cut for a consumer that is done with a batch when it returns (comm.cu: several GPUs). *host holds
// the pinned buffers, created on first use and kept by the caller until stream_host_free
int stream_fasta_chunks(const char* path, void** host,
                        const std::function<int(const char* text, uint64_t len, const uint64_t* beg, const uint64_t* end, uint64_t n_rec)>& on_batch);
void stream_host_free(void* host);
// host_query.cu: the contexts of the host-buffer entry points, created on first use and kept with the index
void host_pool_free(void* pool);
// a batch of records of a text in host memory ([beg[i], end[i]), beg has n+1 entries; end == null: end[i] = beg[i+1]) through
// the read kernels: H2D (chunked, overlapped, optionally 2-bit packed on the host), kernels, D2H of the counters (and ids)
int host_query_records(const struct ::blight_index* idx, const char* text, uint64_t len, const uint64_t* beg, const uint64_t* end,
                       uint64_t n, const uint64_t* koff, int64_t* ids_out, uint64_t total_kmers, uint64_t* ctr, bool allow_pack);
}  // namespace blight
