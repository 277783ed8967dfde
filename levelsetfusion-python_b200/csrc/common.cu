// common.cu -- error reporting, stream-ordered scratch memory, host<->device staging.
#include "common.cuh"

#include <cstring>
#include <atomic>
#include <mutex>

namespace lsf {

static thread_local std::string g_last_error;
static std::atomic<long long> g_launch_count { 0 };

void count_launches(int n) {
	g_launch_count.fetch_add(n, std::memory_order_relaxed);
}

void set_error(const char* fmt, ...) {
	char buffer[1024];
	va_list args;
	va_start(args, fmt);
	vsnprintf(buffer, sizeof(buffer), fmt, args);
	va_end(args);
	g_last_error = buffer;
	// the reference prints assertion messages to stderr as well (error_handling/throw_assert.hpp:67-74)
	fprintf(stderr, "[lsf_b200] %s\n", buffer);
}

static void configure_pool_once() {
	static std::once_flag flags[64];
	int device = 0;
	if (cudaGetDevice(&device) != cudaSuccess || device < 0 || device >= 64) return;
	std::call_once(flags[device], [device]() {
		cudaMemPool_t pool;
		if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
			unsigned long long threshold = ~0ull;  // keep freed blocks cached for the next optimize() call
			cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
		}
	});
}

Arena::Arena(cudaStream_t stream) : stream_(stream) {
	configure_pool_once();
}

Arena::~Arena() {
	for (auto it = blocks_.rbegin(); it != blocks_.rend(); ++it) cudaFreeAsync(*it, stream_);
}

int Arena::alloc_bytes(void** out, size_t bytes) {
	*out = nullptr;
	if (bytes == 0) bytes = 16;
	void* p = nullptr;
	LSF_CUDA(cudaMallocAsync(&p, bytes, stream_));
	blocks_.push_back(p);
	total_ += bytes;
	*out = p;
	return LSF_OK;
}

int to_device(Arena& arena, const float* src, size_t count, int memory_kind, cudaStream_t stream, const float** out) {
	if (memory_kind == LSF_DEVICE) {
		*out = src;
		return LSF_OK;
	}
	float* staged = nullptr;
	LSF_TRY(arena.alloc(&staged, count));
	LSF_CUDA(cudaMemcpyAsync(staged, src, count * sizeof(float), cudaMemcpyHostToDevice, stream));
	*out = staged;
	return LSF_OK;
}

int from_device(const float* src_dev, float* dst, size_t count, int memory_kind, cudaStream_t stream) {
	if (memory_kind == LSF_DEVICE) {
		if (src_dev != dst)
			LSF_CUDA(cudaMemcpyAsync(dst, src_dev, count * sizeof(float), cudaMemcpyDeviceToDevice, stream));
		return LSF_OK;
	}
	LSF_CUDA(cudaMemcpyAsync(dst, src_dev, count * sizeof(float), cudaMemcpyDeviceToHost, stream));
	LSF_CUDA(cudaStreamSynchronize(stream));
	return LSF_OK;
}

int make_taps(const float* kernel_host, int kernel_size, Taps* taps) {
	std::memset(taps, 0, sizeof(Taps));
	LSF_REQUIRE(kernel_host != nullptr && kernel_size > 0, "convolution kernel is empty");
	LSF_REQUIRE(kernel_size <= LSF_MAX_KERNEL_SIZE, "convolution kernel has %d taps, at most %d are supported",
			kernel_size, LSF_MAX_KERNEL_SIZE);
	// the reference's ring buffer (convolution.cpp:83-115) only works for odd sizes
	LSF_REQUIRE(kernel_size % 2 == 1, "convolution kernel size must be odd, got %d", kernel_size);
	taps->size = kernel_size;
	taps->radius = kernel_size / 2;
	// flip, see reference convolution.cpp:232-235: out[i] = sum_j in[i - r + j] * kernel[K - 1 - j]
	for (int j = 0; j < kernel_size; j++) taps->k[j] = kernel_host[kernel_size - 1 - j];
	return LSF_OK;
}

}  // namespace lsf

extern "C" const char* lsf_last_error(void) {
	return lsf::g_last_error.c_str();
}

extern "C" long long lsf_launch_count(void) {
	return lsf::g_launch_count.load();
}

extern "C" int lsf_version(void) {
	return 100;  // 0.1.0
}
