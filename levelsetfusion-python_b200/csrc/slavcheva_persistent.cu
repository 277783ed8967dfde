// slavcheva_persistent.cu -- SobolevFusion / KillingFusion on SMALL fields (the reference's 2D experiments, BASELINE.json
// configs[0]: 128 x 128): a whole polling chunk of iterations in ONE cooperative launch.
//
// A 2D field of 16 K voxels keeps 64 blocks busy for a microsecond per kernel: with five launches per iteration the
// optimizer (reference SobolevOptimizer2d::optimize, cpp/src/nonrigid_optimization/slavcheva/sobolev_optimizer2d.cpp:71-138;
// SlavchevaOptimizer2d.optimize, nonrigid_opt/slavcheva/slavcheva_optimizer2d.py:332-408) is bound by launch latency
// (1.6 ms for 52 iterations at 128 x 128). Here the grid stays resident: per iteration the phases of the generic kernels --
// gradient terms | filter passes | re-warp + maximum warp length | termination test -- run back to back, separated by grid-wide
// barriers (cooperative groups) only where a phase reads other threads' results (1 + passes barriers per iteration), and read what the kernels would have been launched with from a record per iteration
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

template<int D>
__global__ void __launch_bounds__(PERSISTENT_THREADS) k_slav_persistent(const SlavIterationCommand* commands, int count,
		SlavParams p, int N, const unsigned* max_sq_bits, int* status, int first_iteration, int max_iterations) {
	cg::grid_group grid = cg::this_grid();
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
		// so the last pass and the re-warp of a voxel run back to back in the thread that owns it; without a Sobolev kernel
		// the same holds for the gradient terms and the re-warp.
		for (int idx = tid; idx < N; idx += stride) slav_gradient_at<D>(command.gradient, idx);
		for (int pass = 0; pass + 1 < command.passes; pass++) {
			grid.sync();
			for (int idx = tid; idx < N; idx += stride) slav_filter_axis_at<D>(command.pass[pass], idx);
		}
		if (command.passes > 0) grid.sync();
		float sq_report = 0.0f;
		for (int idx = tid; idx < N; idx += stride) {
			if (command.passes > 0) slav_filter_axis_at<D>(command.pass[command.passes - 1], idx);
			slav_resample_at<D>(command.resample, idx, sq_report);
		}
		if (command.resample.max_sq_bits != nullptr) block_atomic_max(sq_report, command.resample.max_sq_bits);
		grid.sync();
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
		if (cooperative && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_slav_persistent<2>, PERSISTENT_THREADS, 0)
				== cudaSuccess)
			blocks = sms * per_sm;
		if (blocks <= 0) blocks = -1;
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
	const unsigned blocks = (unsigned) std::min<long long>(div_up(N, PERSISTENT_THREADS), resident_blocks());
	int n = (int) N;
	SlavParams params = p;
	void* arguments[] = { (void*) &commands_dev, (void*) &count, (void*) &params, (void*) &n, (void*) &max_sq_bits, (void*) &status,
			(void*) &first_iteration, (void*) &max_iterations };
	LSF_CUDA(cudaLaunchCooperativeKernel((const void*) k_slav_persistent<2>, dim3(counted(blocks)), dim3(PERSISTENT_THREADS), arguments,
			0, stream));
	return LSF_OK;
}

}  // namespace lsf
