// common.cu -- error reporting, stream-ordered scratch memory, host<->device staging.
#include "common.cuh"

#include <algorithm>
#include <cstring>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <unistd.h>
#include <mutex>
#include <thread>
#include <vector>

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

// ---------------------------------------------------------------------------------------------- host-side trace
// LSF_TRACE=1: host time between labelled points of a call, printed to stderr at the point named "end" (where does a
// launch-bound call spend its time: allocation, staging, enqueueing or waiting)
void trace_point(const char* label) {
	static const bool enabled = []() {
		const char* e = getenv("LSF_TRACE");
		return e && e[0] == '1';
	}();
	if (!enabled) return;
	static thread_local std::vector<std::pair<const char*, std::chrono::steady_clock::time_point>> points;
	points.emplace_back(label, std::chrono::steady_clock::now());
	if (std::strcmp(label, "end") != 0) return;
	fprintf(stderr, "[lsf_b200 trace]");
	for (size_t i = 1; i < points.size(); i++)
		fprintf(stderr, " %s %.1f us |", points[i].first,
				std::chrono::duration<double, std::micro>(points[i].second - points[i - 1].second).count());
	fprintf(stderr, " total %.1f us\n", std::chrono::duration<double, std::micro>(points.back().second - points.front().second).count());
	points.clear();
}

// ---------------------------------------------------------------------------------------------- scratch memory
// The library allocates from a pool of its own (one per device), not from the device's default pool: the host
// application's allocator state is left alone. Freed blocks stay cached for the next call (one 256^3 optimize() needs
// 2.1 GiB, a batch of 64 pairs of 128^3 10 GiB; re-allocating them per call costs more than the optimisation);
// LSF_POOL_KEEP_MB bounds the cache, lsf_trim() returns everything that is not in use.
static cudaMemPool_t g_pools[64] = { nullptr };

static cudaMemPool_t device_pool() {
	static std::once_flag flags[64];
	int device = 0;
	if (cudaGetDevice(&device) != cudaSuccess || device < 0 || device >= 64) return nullptr;
	std::call_once(flags[device], [device]() {
		cudaMemPoolProps props;
		std::memset(&props, 0, sizeof(props));
		props.allocType = cudaMemAllocationTypePinned;
		props.handleTypes = cudaMemHandleTypeNone;
		props.location.type = cudaMemLocationTypeDevice;
		props.location.id = device;
		cudaMemPool_t pool = nullptr;
		if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) {
			cudaGetLastError();
			return;  // fall back to the default pool (alloc_bytes)
		}
		const char* keep = getenv("LSF_POOL_KEEP_MB");
		unsigned long long threshold = keep && atoll(keep) >= 0 ? (unsigned long long) atoll(keep) << 20 : ~0ull;
		cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
		g_pools[device] = pool;
	});
	return g_pools[device];
}

Arena::Arena(cudaStream_t stream) : stream_(stream) {
	device_pool();
}

Arena::~Arena() {
	for (auto it = blocks_.rbegin(); it != blocks_.rend(); ++it) cudaFreeAsync(*it, stream_);
}

int Arena::alloc_bytes(void** out, size_t bytes) {
	*out = nullptr;
	if (bytes == 0) bytes = 16;
	void* p = nullptr;
	cudaMemPool_t pool = device_pool();
	if (pool != nullptr) LSF_CUDA(cudaMallocFromPoolAsync(&p, bytes, pool, stream_));
	else LSF_CUDA(cudaMallocAsync(&p, bytes, stream_));
	blocks_.push_back(p);
	total_ += bytes;
	*out = p;
	return LSF_OK;
}

// ---------------------------------------------------------------------------------------------- host <-> device staging
// numpy arrays are pageable memory: a cudaMemcpyAsync from / to them is staged by the driver through a small bounce
// buffer at 6 GB/s (measured: 52 ms for the 335 MB of a 256^3 optimize()). The library keeps a ring of two pinned
// chunks per thread instead: worker threads copy chunk k + 1 between the caller's array and the ring while the DMA
// engine moves chunk k. Pinned or registered host memory (cudaPointerGetAttributes) goes straight to the DMA engine.
namespace {

constexpr size_t RING_CHUNK = 32u << 20;

struct StagingRing {
	unsigned char* chunk[2] = { nullptr, nullptr };
	cudaEvent_t done[2] = { nullptr, nullptr };
	bool ok = false;
	StagingRing() {
		ok = cudaHostAlloc(reinterpret_cast<void**>(&chunk[0]), RING_CHUNK, cudaHostAllocDefault) == cudaSuccess
				&& cudaHostAlloc(reinterpret_cast<void**>(&chunk[1]), RING_CHUNK, cudaHostAllocDefault) == cudaSuccess
				&& cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming) == cudaSuccess
				&& cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming) == cudaSuccess;
		if (!ok) cudaGetLastError();
	}
	~StagingRing() {
		// process exit: the CUDA context may be gone already, nothing to release explicitly
	}
};

StagingRing& staging_ring() {
	static thread_local StagingRing ring;
	return ring;
}

bool is_pageable(const void* host) {
	cudaPointerAttributes attributes;
	if (cudaPointerGetAttributes(&attributes, host) != cudaSuccess) {
		cudaGetLastError();
		return true;
	}
	return attributes.type == cudaMemoryTypeUnregistered;
}

// memcpy on several threads (first-touch page faults of a fresh destination array are spread over the threads too).
// The workers are started once and sleep between copies: spawning threads per 32 MB chunk cost ~1 ms per optimize().
class CopyPool {
public:
	static CopyPool& instance() {
		static CopyPool* pool = new CopyPool();  // never destroyed: the workers may outlive static destruction
		return *pool;
	}
	// the worker threads live in the process that created the pool: a forked child copies on its own thread
	unsigned workers() const {
		return getpid() == owner_ ? worker_count_ : 1;
	}
	// copies [0, bytes) in `worker_count_` slices; the caller takes slice 0
	void copy(void* dst, const void* src, size_t bytes) {
		std::lock_guard<std::mutex> serialise(callers_);  // one copy at a time (callers of different threads queue up)
		const size_t slice = ((bytes / worker_count_) + 4095) & ~(size_t) 4095;
		{
			std::lock_guard<std::mutex> lock(mutex_);
			dst_ = static_cast<unsigned char*>(dst);
			src_ = static_cast<const unsigned char*>(src);
			bytes_ = bytes;
			slice_ = slice;
			pending_ = worker_count_ - 1;
			generation_++;
		}
		wake_.notify_all();
		std::memcpy(dst, src, std::min(slice, bytes));
		std::unique_lock<std::mutex> lock(mutex_);
		done_.wait(lock, [&]() { return pending_ == 0; });
	}

private:
	CopyPool() {
		unsigned cores = std::max(1u, std::thread::hardware_concurrency());
		// with several ranks on the host (torchrun exports LOCAL_WORLD_SIZE) each rank takes its share of the cores -- all
		// ranks stage their inputs at the same time; LSF_COPY_THREADS overrides
		const char* local_world = getenv("LOCAL_WORLD_SIZE");
		if (local_world && atoi(local_world) > 1) cores = std::max(1u, cores / (unsigned) atoi(local_world));
		worker_count_ = std::min(8u, cores);
		const char* fixed = getenv("LSF_COPY_THREADS");
		if (fixed && atoi(fixed) >= 1) worker_count_ = std::min(64u, (unsigned) atoi(fixed));
		for (unsigned w = 1; w < worker_count_; w++) std::thread([this, w]() { run(w); }).detach();
	}
	void run(unsigned w) {
		unsigned long long seen = 0;
		for (;;) {
			std::unique_lock<std::mutex> lock(mutex_);
			wake_.wait(lock, [&]() { return generation_ != seen; });
			seen = generation_;
			unsigned char* dst = dst_;
			const unsigned char* src = src_;
			const size_t begin = w * slice_, bytes = bytes_, slice = slice_;
			lock.unlock();
			if (begin < bytes) std::memcpy(dst + begin, src + begin, std::min(slice, bytes - begin));
			lock.lock();
			if (--pending_ == 0) done_.notify_one();
		}
	}
	std::mutex callers_, mutex_;
	std::condition_variable wake_, done_;
	unsigned worker_count_ = 1;
	const pid_t owner_ = getpid();
	unsigned char* dst_ = nullptr;
	const unsigned char* src_ = nullptr;
	size_t bytes_ = 0, slice_ = 0;
	unsigned pending_ = 0;
	unsigned long long generation_ = 0;
};

void parallel_copy(void* dst, const void* src, size_t bytes) {
	CopyPool& pool = CopyPool::instance();
	if (bytes < (4u << 20) || pool.workers() == 1) {
		std::memcpy(dst, src, bytes);
		return;
	}
	pool.copy(dst, src, bytes);
}

}  // namespace

int copy_host_to_device(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t stream) {
	StagingRing& ring = staging_ring();
	if (!is_pageable(src_host) || !ring.ok || bytes < (1u << 20)) {
		LSF_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, stream));
		return LSF_OK;
	}
	size_t offset = 0;
	for (int k = 0; offset < bytes; k++, offset += RING_CHUNK) {
		const int slot = k & 1;
		const size_t length = std::min(RING_CHUNK, bytes - offset);
		if (k >= 2) LSF_CUDA(cudaEventSynchronize(ring.done[slot]));  // the DMA engine has drained this chunk
		parallel_copy(ring.chunk[slot], static_cast<const unsigned char*>(src_host) + offset, length);
		LSF_CUDA(cudaMemcpyAsync(static_cast<unsigned char*>(dst_dev) + offset, ring.chunk[slot], length, cudaMemcpyHostToDevice,
				stream));
		LSF_CUDA(cudaEventRecord(ring.done[slot], stream));
	}
	// the ring is re-used by the next copy of this thread: its chunks must have left
	LSF_CUDA(cudaEventSynchronize(ring.done[0]));
	LSF_CUDA(cudaEventSynchronize(ring.done[1]));
	return LSF_OK;
}

// device -> host; returns after the data has arrived
int copy_device_to_host(void* dst_host, const void* src_dev, size_t bytes, cudaStream_t stream) {
	StagingRing& ring = staging_ring();
	if (!is_pageable(dst_host) || !ring.ok || bytes < (1u << 20)) {
		LSF_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, stream));
		LSF_CUDA(cudaStreamSynchronize(stream));
		return LSF_OK;
	}
	const int chunks = (int) ((bytes + RING_CHUNK - 1) / RING_CHUNK);
	auto request = [&](int k) {
		const size_t offset = (size_t) k * RING_CHUNK;
		cudaMemcpyAsync(ring.chunk[k & 1], static_cast<const unsigned char*>(src_dev) + offset, std::min(RING_CHUNK, bytes - offset),
				cudaMemcpyDeviceToHost, stream);
		cudaEventRecord(ring.done[k & 1], stream);
	};
	request(0);
	for (int k = 0; k < chunks; k++) {
		if (k + 1 < chunks) request(k + 1);  // the other chunk: its previous contents were copied out in the last trip
		LSF_CUDA(cudaEventSynchronize(ring.done[k & 1]));
		const size_t offset = (size_t) k * RING_CHUNK;
		parallel_copy(static_cast<unsigned char*>(dst_host) + offset, ring.chunk[k & 1], std::min(RING_CHUNK, bytes - offset));
	}
	LSF_CUDA(cudaGetLastError());
	return LSF_OK;
}

int to_device(Arena& arena, const float* src, size_t count, int memory_kind, cudaStream_t stream, const float** out) {
	if (memory_kind == LSF_DEVICE) {
		*out = src;
		return LSF_OK;
	}
	float* staged = nullptr;
	LSF_TRY(arena.alloc(&staged, count));
	LSF_TRY(copy_host_to_device(staged, src, count * sizeof(float), stream));
	*out = staged;
	return LSF_OK;
}

int from_device(const float* src_dev, float* dst, size_t count, int memory_kind, cudaStream_t stream) {
	if (memory_kind == LSF_DEVICE) {
		if (src_dev != dst)
			LSF_CUDA(cudaMemcpyAsync(dst, src_dev, count * sizeof(float), cudaMemcpyDeviceToDevice, stream));
		return LSF_OK;
	}
	return copy_device_to_host(dst, src_dev, count * sizeof(float), stream);
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

extern "C" int lsf_trim(void) {
	int device = 0;
	LSF_CUDA(cudaGetDevice(&device));
	LSF_CUDA(cudaDeviceSynchronize());
	if (device >= 0 && device < 64 && lsf::g_pools[device] != nullptr) LSF_CUDA(cudaMemPoolTrimTo(lsf::g_pools[device], 0));
	return LSF_OK;
}

extern "C" int lsf_version(void) {
	return 100;  // 0.1.0
}
