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

template<bool TIKHONOV>
__global__ void __launch_bounds__(THREADS) k_hier_level2d(HierIterArgs2 a, ConvArgs2 c, int use_kernel, float* g_post,
		float* scratch, int first_iteration, int count) {
	cg::grid_group grid = cg::this_grid();
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
			grid.sync();
			c.in = scratch;
			c.out = g_post;
			for (long long idx = tid; idx < g.N; idx += stride) {
				float unused = 0.0f;
				convolve_axis2d_at<0, false>(c, (int) (idx / g.W), (int) (idx % g.W), idx, unused);
			}
			grid.sync();
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
		grid.sync();
	}
}

int resident_blocks() {
	static int blocks = 0;
	if (blocks == 0) {
		int device = 0, sms = 0, per_sm_a = 0, per_sm_b = 0, cooperative = 0;
		cudaGetDevice(&device);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
		cudaDeviceGetAttribute(&cooperative, cudaDevAttrCooperativeLaunch, device);
		if (cooperative && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_a, k_hier_level2d<true>, THREADS, 0) == cudaSuccess
				&& cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_b, k_hier_level2d<false>, THREADS, 0) == cudaSuccess)
			blocks = sms * std::min(per_sm_a, per_sm_b);
		if (blocks <= 0) blocks = -1;
	}
	return blocks;
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
	const unsigned blocks = (unsigned) std::min<long long>(div_up(N, THREADS), resident_blocks());
	HierIterArgs2 a = gradient;
	ConvArgs2 c = filter;
	int kernel_flag = use_kernel ? 1 : 0;
	void* arguments[] = { (void*) &a, (void*) &c, (void*) &kernel_flag, (void*) &g_post, (void*) &scratch, (void*) &first_iteration,
			(void*) &count };
	const void* kernel = tikhonov ? (const void*) k_hier_level2d<true> : (const void*) k_hier_level2d<false>;
	LSF_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(counted(blocks)), dim3(THREADS), arguments, 0, stream));
	return LSF_OK;
}

}  // namespace lsf
