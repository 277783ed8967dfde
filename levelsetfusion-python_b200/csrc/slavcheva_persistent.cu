// slavcheva_persistent.cu -- SobolevFusion / KillingFusion on SMALL fields (the reference's 2D experiments, BASELINE.json
// configs[0]: 128 x 128): a whole polling chunk of iterations in ONE cooperative launch.
//
// A 2D field of 16 K voxels keeps 64 blocks busy for a microsecond per kernel: with five launches per iteration the
// optimizer (reference SobolevOptimizer2d::optimize, cpp/src/nonrigid_optimization/slavcheva/sobolev_optimizer2d.cpp:71-138;
// SlavchevaOptimizer2d.optimize, nonrigid_opt/slavcheva/slavcheva_optimizer2d.py:332-408) is bound by launch latency
// (1.6 ms for 52 iterations at 128 x 128). Here the grid stays resident: per iteration the phases of the generic kernels --
// gradient terms | filter passes | re-warp + maximum warp length | termination test -- run back to back, separated by grid-wide
// barriers only where a phase reads or overwrites other threads' data (1 + max(passes, 1) barriers per iteration), and read what the kernels would have been launched with from a record per iteration
// (SlavIterationCommand, written by the same host code that otherwise launches the kernels: buffer rotation and all
// semantics switches stay in one place). The per-voxel code is that of the kernels (slav_gradient_at, slav_filter_axis_at,
// slav_resample_at): results are bit-identical.
//
// Every field is written in one phase and read in the next one of the SAME launch, so nothing here may travel through the
// read-only (non-coherent) data path: __ldg is mapped to a plain load for this translation unit.
#include "common.cuh"

#include <cooperative_groups.h>

#define __ldg(pointer) (*(pointer))
#include "slavcheva.cuh"

namespace cg = cooperative_groups;

namespace lsf {
namespace {

constexpr int PERSISTENT_THREADS = 256;
// Fields of up to 16 K voxels (128 x 128, BASELINE.json configs[0]) run in ONE thread-block cluster of up to 16 blocks: the
// hardware cluster barrier (~0.2 us) replaces the grid barrier through L2 (~1 us), which is most of such an iteration.
constexpr int CLUSTER_THREADS = 1024, CLUSTER_BLOCKS = 16;

// barrier between two phases: grid-wide (cooperative launch) or cluster-wide (the whole grid is one cluster)
template<bool CLUSTER>
__device__ __forceinline__ void phase_barrier() {
	if (CLUSTER) cg::this_cluster().sync();
	else cg::this_grid().sync();
}

template<int D, bool CLUSTER>
__global__ void __launch_bounds__(CLUSTER ? CLUSTER_THREADS : PERSISTENT_THREADS) k_slav_persistent(const SlavIterationCommand* commands,
		int count, SlavParams p, int N, const unsigned* max_sq_bits, int* status, int first_iteration, int max_iterations) {
	__shared__ SlavIterationCommand command;
	static_assert(sizeof(SlavIterationCommand) % 4 == 0, "copied word by word");
	const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
	// status[first_iteration] comes from the previous launch; later decisions are taken by every thread itself
	bool finished = *reinterpret_cast<volatile int*>(status + first_iteration) != 0;
	for (int j = 0; j < count; j++) {
		const int it = first_iteration + j;
		if (finished) {
			if (tid == 0)
				for (int k = j; k < count; k++) status[first_iteration + k + 1] = 1;  // sticky, like k_slav_decide
			break;
		}
		__syncthreads();  // the previous iteration's readers of `command` are done
		for (int w = threadIdx.x; w < (int) (sizeof(SlavIterationCommand) / 4); w += blockDim.x)
			reinterpret_cast<int*>(&command)[w] = reinterpret_cast<const int*>(commands + j)[w];
		__syncthreads();
		// Grid barriers stand only where a phase reads what OTHER threads wrote in the phase before: a filter pass reads its
		// input along the pass axis. The re-warp needs the filtered update of its own voxel only (and the old live field),
		// so the last pass and the re-warp of a voxel run back to back in the thread that owns it.
		for (int idx = tid; idx < N; idx += stride) slav_gradient_at<D>(command.gradient, idx);
		for (int pass = 0; pass + 1 < command.passes; pass++) {
			phase_barrier<CLUSTER>();
			for (int idx = tid; idx < N; idx += stride) slav_filter_axis_at<D>(command.pass[pass], idx);
		}
		// also without a filter: the re-warp stores the new warp vector of its voxel, which the smoothing terms of the
		// neighbouring voxels (gradient phase of this iteration, other threads) still read
		phase_barrier<CLUSTER>();
		float sq_report = 0.0f;
		for (int idx = tid; idx < N; idx += stride) {
			if (command.passes > 0) slav_filter_axis_at<D>(command.pass[command.passes - 1], idx);
			slav_resample_at<D>(command.resample, idx, sq_report);
		}
		if (command.resample.max_sq_bits != nullptr) block_atomic_max(sq_report, command.resample.max_sq_bits);
		phase_barrier<CLUSTER>();
		// k_slav_decide, evaluated by every thread (one barrier less); thread 0 records it for the host
		const float max_warp = sqrtf(__uint_as_float(*reinterpret_cast<const volatile unsigned*>(max_sq_bits + it)));
		finished = slav_finished(p, it + 1, max_iterations, max_warp);
		if (tid == 0) status[it + 1] = finished ? 1 : 0;
	}
}

int resident_blocks() {
	static int blocks = 0;
	if (blocks == 0) {
		int device = 0, sms = 0, per_sm = 0;
		cudaGetDevice(&device);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
		int cooperative = 0;
		cudaDeviceGetAttribute(&cooperative, cudaDevAttrCooperativeLaunch, device);
		if (cooperative && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_slav_persistent<2, false>, PERSISTENT_THREADS, 0)
				== cudaSuccess)
			blocks = sms * per_sm;
		if (blocks <= 0) blocks = -1;
	}
	return blocks;
}

// largest cluster (16, else 8 blocks of CLUSTER_THREADS threads) the device schedules; 0 = none (LSF_SLAV_CLUSTER=0 keeps the
// cooperative grid: A/B tests)
int cluster_blocks() {
	const char* env = getenv("LSF_SLAV_CLUSTER");
	if (env && env[0] == '0') return 0;
	static int blocks = -1;
	if (blocks < 0) {
		blocks = 0;
		const void* kernel = (const void*) k_slav_persistent<2, true>;
		for (int size = CLUSTER_BLOCKS; size >= 8 && blocks == 0; size /= 2) {
			bool ok = size <= 8 || cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
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
			if (ok && cudaOccupancyMaxActiveClusters(&clusters, kernel, &config) == cudaSuccess && clusters >= 1) blocks = size;
		}
		cudaGetLastError();  // a refused attribute / query is not an error of the caller
	}
	return blocks;
}

}  // namespace

long long slav_persistent_capacity() {
	const int blocks = resident_blocks();
	// one voxel per thread up to a full wave; beyond that the ordinary kernels are no longer launch-bound
	return blocks > 0 ? (long long) blocks * PERSISTENT_THREADS : 0;
}

int launch_slav_persistent2d(const SlavIterationCommand* commands_dev, int count, const SlavParams& p, long long N,
		const unsigned* max_sq_bits, int* status, int first_iteration, int max_iterations, cudaStream_t stream) {
	LSF_REQUIRE(N > 0 && N <= slav_persistent_capacity(), "field of %lld voxels does not fit the single-launch path", N);
	int n = (int) N;
	SlavParams params = p;
	if (cluster_blocks() > 0 && N <= (long long) cluster_blocks() * CLUSTER_THREADS) {
		// one voxel per thread where the cluster has the threads; small fields spread over all SMs of the cluster
		unsigned threads = 128;
		while (threads < (unsigned) CLUSTER_THREADS && (long long) threads * cluster_blocks() < N) threads *= 2;
		const unsigned blocks = (unsigned) std::min<long long>(cluster_blocks(), div_up(N, (long long) threads));
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
		LSF_CUDA(cudaLaunchKernelEx(&config, k_slav_persistent<2, true>, commands_dev, count, params, n, max_sq_bits, status, first_iteration,
				max_iterations));
		return LSF_OK;
	}
	const unsigned blocks = (unsigned) std::min<long long>(div_up(N, PERSISTENT_THREADS), resident_blocks());
	void* arguments[] = { (void*) &commands_dev, (void*) &count, (void*) &params, (void*) &n, (void*) &max_sq_bits, (void*) &status,
			(void*) &first_iteration, (void*) &max_iterations };
	LSF_CUDA(cudaLaunchCooperativeKernel((const void*) k_slav_persistent<2, false>, dim3(counted(blocks)), dim3(PERSISTENT_THREADS), arguments,
			0, stream));
	return LSF_OK;
}

}  // namespace lsf
