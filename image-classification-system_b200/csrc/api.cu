// Library-level entry points of libb2ingest: version, thread-local error text, device check.
#include "common.cuh"

#include <cstdlib>
#include <mutex>

namespace b2 {

char *error_buffer() {
    static thread_local char buf[512] = "";
    return buf;
}

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

int sm_count() {
    static int cached[64];
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    std::lock_guard<std::mutex> lock(mu);
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace b2

// The streaming ingest keeps tens of kernels (one hash launch per ~4 GiB of messages, each alive for as long as its
// longest message hashes) and several copy / resize streams in flight.  Streams share hardware work queues
// ("connections", 8 by default, 32 at most) and the GPU overlaps only a few kernels per queue: measured on the
// config-3 stream, 1 connection = 9.5 GB/s, 8 = 25.6 GB/s, 32 = PCIe bound.  Ask for 32 unless the user chose; the
// variable is read when the CUDA context is created, so this only helps when the library is loaded before that.
__attribute__((constructor)) static void b2_on_load() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

extern "C" int b2_version(void) { return 2; }

extern "C" const char *b2_last_error(void) { return b2::error_buffer(); }

extern "C" int b2_init(int device) {
    using namespace b2;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(B2_ERR_NO_DEVICE, "no CUDA device visible (%s); libb2ingest has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    B2_REQUIRE(device >= 0 && device < count, "b2_init: device %d out of range (0..%d)", device, count - 1);
    int major = 0, minor = 0;
    B2_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    B2_CUDA_CHECK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    if (major != 10)
        return fail(B2_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                    device, major, minor);
    B2_CUDA_CHECK(cudaSetDevice(device));
    return B2_OK;
}

extern "C" int b2_device_sm_count(int device, int *sm_count_out) {
    using namespace b2;
    B2_REQUIRE(sm_count_out != nullptr, "b2_device_sm_count: null output");
    B2_CUDA_CHECK(cudaDeviceGetAttribute(sm_count_out, cudaDevAttrMultiProcessorCount, device));
    return B2_OK;
}
