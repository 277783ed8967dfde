// hier2d_persistent.cu -- 2D hierarchical optimizer on small fields: ALL iterations of a pyramid level in one cooperative
// launch.
//
// The reference's 2D experiments and tests run Optimizer<MatrixXf,MatrixXv2f> (cpp/src/nonrigid_optimization/hierarchical/
// optimizer.tpp:134-212) on 16 x 16 ... 512 x 512 fields: a level is 100 iterations of one to three kernels that each take
// a few microseconds -- launch latency, not bandwidth, sets the time (20 us per iteration with a Sobolev kernel). Here the
// grid stays resident for the whole level: gradient stage | axis-0 pass | axis-1 pass + warp update + max norm run back to
// back, separated by grid-wide barriers (cooperative groups), with the termination test of optimizer.tpp:166-171 evaluated by
// every thread at the head of an iteration from the previous iteration's maximum. Per-pixel code = the bodies of
// k_hier_gradient2d / k_convolve_axis2d (kernels2d.cuh): bit-identical results.
//
// The filter passes read what other threads wrote in the phase before of the SAME launch: nothing here may use the
// read-only (non-coherent) data path, so __ldg is mapped to a plain load for this translation unit.
#include "common.cuh"

#include <cooperative_groups.h>

#define __ldg(pointer) (*(pointer))
#include "kernels2d.cuh"

#include <algorithm>

namespace cg = cooperative_groups;

namespace lsf {
namespace {

constexpr int THREADS = 256;
// Fields of up to 16 K pixels (128 x 128, the reference's 2D experiment size) run in ONE thread-block cluster of up to 16
// blocks instead of a cooperative grid: the hardware cluster barrier costs ~0.2 us where the grid barrier (atomics through
// L2) costs ~1 us, and with three barriers per iteration the barrier is most of a launch-bound iteration.
constexpr int CLUSTER_THREADS = 1024, CLUSTER_BLOCKS = 16;

// the barrier between two phases: grid-wide (cooperative launch) or cluster-wide (the whole grid is one cluster); both order
// the global-memory writes of the phase before against the reads of the phase after
template<bool CLUSTER>
__device__ __forceinline__ void phase_barrier() {
	if (CLUSTER) cg::this_cluster().sync();
	else cg::this_grid().sync();
}

template<bool TIKHONOV, bool CLUSTER>
__global__ void __launch_bounds__(CLUSTER ? CLUSTER_THREADS : THREADS, 1) k_hier_level2d(HierIterArgs2 a, ConvArgs2 c, int use_kernel,
		float* g_post, float* scratch, int first_iteration, int count) {
	const Grid2 g = a.g;
	const long long tid = (long long) blockIdx.x * blockDim.x + threadIdx.x, stride = (long long) gridDim.x * blockDim.x;
	for (int it = first_iteration; it < first_iteration + count; it++) {
		// level_converged(): the slot of the previous iteration is complete (grid barrier / previous launch)
		if (it > 0) {
			const float max_norm = sqrtf(__uint_as_float(*reinterpret_cast<const volatile unsigned*>(a.max_sq_bits + it - 1)));
			if (max_norm < a.threshold) break;
		}
		a.g_prev = g_post;
		float sq = 0.0f;
		if (!use_kernel) {
			// the whole iteration is point-wise apart from the Laplacian of the previous gradient (other buffer)
			a.g_out = TIKHONOV ? scratch : nullptr;
			for (long long idx = tid; idx < g.N; idx += stride) {
				float mine = 0.0f;
				hier_gradient2d_at<TIKHONOV, true>(a, (int) (idx / g.W), (int) (idx % g.W), idx, mine);
				sq = fmaxf(sq, mine);
			}
			if (TIKHONOV) {
				float* t = g_post;
				g_post = scratch;
				scratch = t;
			}
		} else {
			// g_pre -> scratch (stage 1), rows pass scratch -> g_post, columns pass g_post -> scratch (+ update)
			a.g_out = scratch;
			for (long long idx = tid; idx < g.N; idx += stride) {
				float unused = 0.0f;
				hier_gradient2d_at<TIKHONOV, false>(a, (int) (idx / g.W), (int) (idx % g.W), idx, unused);
			}
			phase_barrier<CLUSTER>();
			c.in = scratch;
			c.out = g_post;
			for (long long idx = tid; idx < g.N; idx += stride) {
				float unused = 0.0f;
				convolve_axis2d_at<0, false>(c, (int) (idx / g.W), (int) (idx % g.W), idx, unused);
			}
			phase_barrier<CLUSTER>();
			c.in = g_post;
			c.out = scratch;
			for (long long idx = tid; idx < g.N; idx += stride) {
				float mine = 0.0f;
				convolve_axis2d_at<1, true>(c, (int) (idx / g.W), (int) (idx % g.W), idx, mine);
				sq = fmaxf(sq, mine);
			}
			float* t = g_post;
			g_post = scratch;
			scratch = t;
		}
		block_atomic_max(sq, a.max_sq_bits + it);
		phase_barrier<CLUSTER>();
	}
}

int resident_blocks() {
	static int blocks = 0;
	if (blocks == 0) {
		int device = 0, sms = 0, per_sm_a = 0, per_sm_b = 0, cooperative = 0;
		cudaGetDevice(&device);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
		cudaDeviceGetAttribute(&cooperative, cudaDevAttrCooperativeLaunch, device);
		if (cooperative && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_a, k_hier_level2d<true, false>, THREADS, 0) == cudaSuccess
				&& cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_b, k_hier_level2d<false, false>, THREADS, 0) == cudaSuccess)
			blocks = sms * std::min(per_sm_a, per_sm_b);
		if (blocks <= 0) blocks = -1;
	}
	return blocks;
}

// largest cluster (16, else 8 blocks of CLUSTER_THREADS threads) the device schedules for both instantiations; 0 = none
// (LSF_HIER2D_CLUSTER=0 keeps the cooperative grid: A/B tests)
int cluster_blocks() {
	const char* env = getenv("LSF_HIER2D_CLUSTER");
	if (env && env[0] == '0') return 0;
	static int blocks = -1;
	if (blocks < 0) {
		blocks = 0;
		for (int size = CLUSTER_BLOCKS; size >= 8 && blocks == 0; size /= 2) {
			bool ok = true;
			for (const void* kernel : { (const void*) k_hier_level2d<true, true>, (const void*) k_hier_level2d<false, true> }) {
				if (size > 8 && cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) ok = false;
				cudaLaunchConfig_t config = {};
				config.gridDim = dim3(size);
				config.blockDim = dim3(CLUSTER_THREADS);
				cudaLaunchAttribute attribute;
				attribute.id = cudaLaunchAttributeClusterDimension;
				attribute.val.clusterDim.x = size;
				attribute.val.clusterDim.y = attribute.val.clusterDim.z = 1;
				config.attrs = &attribute;
				config.numAttrs = 1;
				int clusters = 0;
				if (!ok || cudaOccupancyMaxActiveClusters(&clusters, kernel, &config) != cudaSuccess || clusters < 1) ok = false;
			}
			if (ok) blocks = size;
		}
		cudaGetLastError();  // a refused attribute / query is not an error of the caller
	}
	return blocks;
}

// blocks x threads of the one-cluster launch for N pixels (one pixel per thread where the cluster has the threads; small
// levels spread over as many SMs as the cluster has, with fewer threads per block)
void cluster_shape(long long N, unsigned& blocks, unsigned& threads) {
	const int most = cluster_blocks();
	threads = 128;
	while (threads < (unsigned) CLUSTER_THREADS && (long long) threads * most < N) threads *= 2;
	blocks = (unsigned) std::min<long long>(most, div_up(N, (long long) threads));
}

}  // namespace

long long hier2d_persistent_capacity() {
	const int blocks = resident_blocks();
	// up to four pixels per thread; larger fields are no longer launch-bound
	return blocks > 0 ? 4ll * blocks * THREADS : 0;
}

int launch_hier2d_persistent(const HierIterArgs2& gradient, const ConvArgs2& filter, bool tikhonov, bool use_kernel,
		float* g_post, float* scratch, int first_iteration, int count, cudaStream_t stream) {
	const long long N = gradient.g.N;
	LSF_REQUIRE(N > 0 && N <= hier2d_persistent_capacity(), "field of %lld pixels does not fit the single-launch path", N);
	HierIterArgs2 a = gradient;
	ConvArgs2 c = filter;
	int kernel_flag = use_kernel ? 1 : 0;
	if (cluster_blocks() > 0 && N <= (long long) cluster_blocks() * CLUSTER_THREADS) {
		unsigned blocks = 0, threads = 0;
		cluster_shape(N, blocks, threads);
		cudaLaunchConfig_t config = {};
		config.gridDim = dim3(counted(blocks));
		config.blockDim = dim3(threads);
		config.stream = stream;
		cudaLaunchAttribute attribute;
		attribute.id = cudaLaunchAttributeClusterDimension;
		attribute.val.clusterDim.x = blocks;
		attribute.val.clusterDim.y = attribute.val.clusterDim.z = 1;
		config.attrs = &attribute;
		config.numAttrs = 1;
		if (tikhonov) LSF_CUDA(cudaLaunchKernelEx(&config, k_hier_level2d<true, true>, a, c, kernel_flag, g_post, scratch, first_iteration, count));
		else LSF_CUDA(cudaLaunchKernelEx(&config, k_hier_level2d<false, true>, a, c, kernel_flag, g_post, scratch, first_iteration, count));
		return LSF_OK;
	}
	const unsigned blocks = (unsigned) std::min<long long>(div_up(N, THREADS), resident_blocks());
	void* arguments[] = { (void*) &a, (void*) &c, (void*) &kernel_flag, (void*) &g_post, (void*) &scratch, (void*) &first_iteration,
			(void*) &count };
	const void* kernel = tikhonov ? (const void*) k_hier_level2d<true, false> : (const void*) k_hier_level2d<false, false>;
	LSF_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(counted(blocks)), dim3(THREADS), arguments, 0, stream));
	return LSF_OK;
}

}  // namespace lsf
