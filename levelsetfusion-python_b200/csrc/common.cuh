// common.cuh -- shared host/device helpers of liblsf_b200.so (sm_100a only).
//
// Numerics contract: every kernel reproduces the reference's float32 operation order without FMA
// contraction (the reference CI build is plain SSE2), so the translation units are
// compiled with --fmad=false and results are bit-identical to the CPU oracle.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <vector>

#include "../../include/lsf_b200.h"

namespace lsf {

// ---------------------------------------------------------------------------------------------- errors
void set_error(const char* fmt, ...);

// LSF_TRACE=1: host time between labelled points of a call (printed at the point named "end")
void trace_point(const char* label);

// Kernel-launch accounting (bench.py's `gpu_launches`): every launch site reports how many kernels it enqueued.
void count_launches(int n);
// wraps the grid argument of every <<<...>>> launch: counts the launch and passes the grid through
inline dim3 counted(dim3 grid) {
	count_launches(1);
	return grid;
}
inline unsigned counted(unsigned grid) {
	count_launches(1);
	return grid;
}

#define LSF_CUDA(call)                                                                              \
	do {                                                                                            \
		cudaError_t lsf_cuda_err__ = (call);                                                        \
		if (lsf_cuda_err__ != cudaSuccess) {                                                        \
			lsf::set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(lsf_cuda_err__), __FILE__, \
					__LINE__, #call);                                                               \
			return LSF_ERR_CUDA;                                                                    \
		}                                                                                           \
	} while (0)

#define LSF_REQUIRE(cond, ...)                   \
	do {                                         \
		if (!(cond)) {                           \
			lsf::set_error(__VA_ARGS__);         \
			return LSF_ERR_INVALID_ARGUMENT;     \
		}                                        \
	} while (0)

#define LSF_TRY(expr)                    \
	do {                                 \
		int lsf_status__ = (expr);       \
		if (lsf_status__ < 0) return lsf_status__; \
	} while (0)

inline bool is_power_of_two(int v) {
	return v > 0 && (v & (v - 1)) == 0;
}

inline unsigned div_up(long long a, long long b) {
	return (unsigned) ((a + b - 1) / b);
}

// ---------------------------------------------------------------------------------------------- device memory
// Stream-ordered scratch allocations (cudaMallocAsync pool, release threshold raised so repeated
// optimize() calls reuse the same blocks). Freed in reverse order when the arena goes out of scope.
class Arena {
public:
	explicit Arena(cudaStream_t stream);
	~Arena();
	template<typename T>
	int alloc(T** out, size_t count) {
		void* p = nullptr;
		int status = alloc_bytes(&p, count * sizeof(T));
		*out = static_cast<T*>(p);
		return status;
	}
	int alloc_bytes(void** out, size_t bytes);
	size_t bytes_allocated() const {
		return total_;
	}
private:
	cudaStream_t stream_;
	std::vector<void*> blocks_;
	size_t total_ = 0;
};

// Stages a host or device input so that kernels see a device pointer.
int to_device(Arena& arena, const float* src, size_t count, int memory_kind, cudaStream_t stream, const float** out);
// Copies a device result to the caller's buffer (host or device).
int from_device(const float* src_dev, float* dst, size_t count, int memory_kind, cudaStream_t stream);

// TSDF generation on device buffers (tsdf.cu); camera_pose is a host 4 x 4 row-major matrix
int tsdf_generate_device(const lsf_tsdf_params* params, const unsigned short* depth_dev, int rows, int cols,
		const float* camera_pose, int image_y_coordinate, int nd, float* field_dev, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------- separable kernel taps
struct Taps {
	float k[LSF_MAX_KERNEL_SIZE];  // flipped: k[j] multiplies in[i - r + j]
	int size;
	int radius;
};
int make_taps(const float* kernel_host, int kernel_size, Taps* taps);

// Planes per block of a kernel that marches along an axis of `extent` planes with `tiles` tiles per plane, each block
// starting with `lead` planes of re-computed halo. Blocks run in waves of `slots` (SMs x resident blocks per SM), so
// the time is about waves x (chunk + lead) plane-steps: pick the split that minimises it. On B200 (444 slots) a 256^3
// level gets 5 chunks of 52 planes -- measured 0.273 ms against 0.304 ms for 4 chunks of 64 (2.3 waves, the third one
// a third full), in the order this model predicts for every split tried (profiles/r1_sweep_chunks_v3.log, profiles/r1_ncu_v4.md).
inline int marching_chunk(int extent, int tiles, int lead, int blocks_per_sm) {
	static int sm_count = 0;
	if (sm_count == 0) {
		int device = 0;
		cudaGetDevice(&device);
		if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sm_count <= 0)
			sm_count = 148;
	}
	const long long slots = (long long) sm_count * blocks_per_sm;
	// without a halo there is nothing to amortise: short marches balance best (measured flat between 10 and 26 planes)
	if (lead == 0) return extent < 16 ? extent : 16;
	int best_chunk = extent;
	long long best_cost = -1;
	for (int parts = 1; parts <= extent; parts++) {
		const int chunk = (extent + parts - 1) / parts;
		if (chunk < 4 && parts > 1) break;
		const long long blocks = (long long) tiles * ((extent + chunk - 1) / chunk);
		const long long cost = ((blocks + slots - 1) / slots) * (chunk + lead);
		if (best_cost < 0 || cost < best_cost) {
			best_cost = cost;
			best_chunk = chunk;
		}
	}
	return best_chunk;
}

// ---------------------------------------------------------------------------------------------- device helpers
#ifdef __CUDACC__

// Programmatic dependent launch (sm_90+): a kernel launched with launch_dependent() may be scheduled while its
// predecessor in the stream still drains -- its blocks run their prologue (barrier set-up, shared-memory clearing)
// and then sleep in pdl_wait() until the predecessor has completed and its writes are visible. Every kernel of the
// iteration loop calls pdl_launch_dependents() first thing (lets the successor in) and pdl_wait() before it touches
// global memory. Both are no-ops for plain launches. LSF_PDL=0 launches without the attribute (A/B).
__device__ __forceinline__ void pdl_launch_dependents() {
	asm volatile("griddepcontrol.launch_dependents;");
}
__device__ __forceinline__ void pdl_wait() {
	asm volatile("griddepcontrol.wait;" ::: "memory");
}

inline bool pdl_enabled() {
	static int enabled = -1;
	if (enabled < 0) {
		const char* e = getenv("LSF_PDL");
		enabled = (e && e[0] == '0') ? 0 : 1;
	}
	return enabled == 1;
}

template<typename... Params, typename... Args>
inline cudaError_t launch_dependent(void (*kernel)(Params...), dim3 grid, dim3 block, size_t shared, cudaStream_t stream,
		Args&&... args) {
	cudaLaunchConfig_t config = {};
	config.gridDim = counted(grid);
	config.blockDim = block;
	config.dynamicSmemBytes = shared;
	config.stream = stream;
	cudaLaunchAttribute attribute;
	attribute.id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attribute.val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
	config.attrs = &attribute;
	config.numAttrs = 1;
	return cudaLaunchKernelEx(&config, kernel, static_cast<Params>(args)...);
}

// Level termination test evaluated on the device so that the host never has to synchronise inside the
// iteration loop: iteration `it` runs iff it == 0 or max||g|| of iteration it-1 is >= threshold
// (reference optimizer.tpp:149,166-171). max_sq_bits[i] holds the bits of max ||g||^2 of iteration i.
__device__ __forceinline__ bool level_converged(const unsigned* __restrict__ max_sq_bits, int iteration,
		float threshold) {
	if (iteration <= 0 || max_sq_bits == nullptr) return false;
	const float max_norm = sqrtf(__uint_as_float(max_sq_bits[iteration - 1]));
	return max_norm < threshold;
}

// Block-wide max of non-negative floats followed by one atomicMax on the float's bit pattern
// (non-negative IEEE floats order like unsigned integers). NaNs are ignored, as in the reference's
// `if (squared_length > max)` (statistics.tpp:65).
__device__ __forceinline__ void block_atomic_max(float value, unsigned* target) {
	__shared__ float warp_max[32];
	if (!(value >= 0.0f)) value = 0.0f;  // NaN -> ignored
#pragma unroll
	for (int offset = 16; offset > 0; offset >>= 1) value = fmaxf(value, __shfl_xor_sync(0xffffffffu, value, offset));
	const int linear = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
	const int lane = linear & 31, warp = linear >> 5;
	const int warps = (blockDim.x * blockDim.y * blockDim.z + 31) >> 5;
	if (lane == 0) warp_max[warp] = value;
	__syncthreads();
	if (warp == 0) {
		value = lane < warps ? warp_max[lane] : 0.0f;
#pragma unroll
		for (int offset = 16; offset > 0; offset >>= 1)
			value = fmaxf(value, __shfl_xor_sync(0xffffffffu, value, offset));
		if (lane == 0 && value > 0.0f) atomicMax(target, __float_as_uint(value));
	}
}

#endif  // __CUDACC__

}  // namespace lsf
