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

#include <cstddef>
#include <map>
#include <mutex>
#include <tuple>

#define __ldg(pointer) (*(pointer))

namespace lsf {
namespace {
// the strip of rows a block of the distributed-shared-memory optimizer (k_slav_strips, below) holds: rows [row_lo, row_hi) of
// every field are in this block's shared memory, row r of the field belongs to block r / rows_per
struct SlavStrip {
	int enabled;  // 0: the fields live in global memory (k_slav_persistent)
	int rank, blocks, rows_per, row_lo, row_hi, W;
};
__shared__ SlavStrip slav_strip;

// a tap of the re-warp's gather: rows outside the block's strip + halo are read from the shared memory of the block that
// owns them (every block lays its fields out the same way, so the slot of a voxel in the owner's shared memory is this
// block's slot shifted by the distance of the two strips)
__device__ __forceinline__ float slav_strip_live_tap(const float* live, int row, int index) {
	if (!slav_strip.enabled || (row >= slav_strip.row_lo && row < slav_strip.row_hi)) return live[index];
	const int owner = min(row / slav_strip.rows_per, slav_strip.blocks - 1);
	const float* slot = live + index + (slav_strip.rank - owner) * slav_strip.rows_per * slav_strip.W;
	return *cooperative_groups::this_cluster().map_shared_rank(slot, owner);
}
}  // namespace
}  // namespace lsf
#define SLAV_LIVE_TAP(live, row, index) lsf::slav_strip_live_tap(live, row, index)
#include "slavcheva.cuh"

namespace cg = cooperative_groups;

namespace lsf {
namespace {

constexpr int PERSISTENT_THREADS = 256;
// Fields of up to 16 K voxels (128 x 128, BASELINE.json configs[0]) run in ONE thread-block cluster of up to 16 blocks: the
// hardware cluster barrier (~0.2 us) replaces the grid barrier through L2 (~1 us), which is most of such an iteration.
constexpr int CLUSTER_THREADS = 1024, CLUSTER_BLOCKS = 16;
constexpr int PHASE_CLOCK_ITERATIONS = 8;

// barrier between two phases: grid-wide (cooperative launch) or cluster-wide (the whole grid is one cluster)
template<bool CLUSTER>
__device__ __forceinline__ void phase_barrier() {
	if (CLUSTER) cg::this_cluster().sync();
	else cg::this_grid().sync();
}

template<int D, bool CLUSTER>
__global__ void __launch_bounds__(CLUSTER ? CLUSTER_THREADS : PERSISTENT_THREADS, 1) k_slav_persistent(const SlavIterationCommand* commands,
		int count, SlavParams p, int N, const unsigned* max_sq_bits, int* status, int first_iteration, int max_iterations,
		unsigned long long* phase_clock) {
	__shared__ SlavIterationCommand command;
	static_assert(sizeof(SlavIterationCommand) % 4 == 0, "copied word by word");
	if (threadIdx.x == 0) slav_strip.enabled = 0;  // the fields are in global memory (read after the barrier below)
	const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
	// status[first_iteration] comes from the previous launch; later decisions are taken by every thread itself
	bool finished = *reinterpret_cast<volatile int*>(status + first_iteration) != 0;
	// LSF_TRACE=2: thread 0 stamps the nanosecond clock at the phase boundaries of the first iterations (8 stamps each)
	const bool stamping = phase_clock != nullptr && tid == 0;
	auto stamp = [&](int j, int slot) {
		if (stamping && j < PHASE_CLOCK_ITERATIONS) {
			unsigned long long now;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
			phase_clock[j * 8 + slot] = now;
		}
	};
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
		stamp(j, 0);
		// Grid barriers stand only where a phase reads what OTHER threads wrote in the phase before: a filter pass reads its
		// input along the pass axis. The re-warp needs the filtered update of its own voxel only (and the old live field),
		// so the last pass and the re-warp of a voxel run back to back in the thread that owns it.
		for (int idx = tid; idx < N; idx += stride) slav_gradient_at<D>(command.gradient, idx);
		stamp(j, 1);
		for (int pass = 0; pass + 1 < command.passes; pass++) {
			phase_barrier<CLUSTER>();
			if (pass == 0) stamp(j, 2);
			for (int idx = tid; idx < N; idx += stride) slav_filter_axis_at<D>(command.pass[pass], idx);
			if (pass == 0) stamp(j, 3);
		}
		// also without a filter: the re-warp stores the new warp vector of its voxel, which the smoothing terms of the
		// neighbouring voxels (gradient phase of this iteration, other threads) still read
		phase_barrier<CLUSTER>();
		stamp(j, 4);
		float sq_report = 0.0f;
		for (int idx = tid; idx < N; idx += stride) {
			if (command.passes > 0) slav_filter_axis_at<D>(command.pass[command.passes - 1], idx);
			slav_resample_at<D>(command.resample, idx, sq_report);
		}
		stamp(j, 5);
		if (command.resample.max_sq_bits != nullptr) block_atomic_max(sq_report, command.resample.max_sq_bits);
		phase_barrier<CLUSTER>();
		stamp(j, 6);
		// k_slav_decide, evaluated by every thread (one barrier less); thread 0 records it for the host
		const float max_warp = sqrtf(__uint_as_float(*reinterpret_cast<const volatile unsigned*>(max_sq_bits + it)));
		finished = slav_finished(p, it + 1, max_iterations, max_warp);
		if (tid == 0) status[it + 1] = finished ? 1 : 0;
		stamp(j, 7);
	}
}


// ---------------------------------------------------------------------------------------------- fields in distributed shared memory
// The phases of k_slav_persistent cost ~1 us each, the barriers between them 2.3 us each (LSF_TRACE=2 phase clock, 128 x 128:
// 13.2 us per iteration): what a barrier waits for is the round trip of the phase's stores to L2 and of the next phase's loads
// from it. k_slav_strips keeps every field of the optimizer in the shared memory of ONE cluster instead: block k owns the rows
// [k * rows_per, (k + 1) * rows_per) of all seven buffers (live x 2, canonical, warp, three update fields) plus `halo` rows on
// either side (halo = filter radius; the term stencils need one row). A phase reads only this block's shared memory, writes
// its own rows and stores the rows its neighbours keep as halo into THEIR shared memory (distributed shared memory); a cluster
// barrier separates the phases; the re-warp's gather reads rows outside strip + halo from their owner's shared memory
// (slav_strip_live_tap); the maximum warp length travels through a slot per block in every block's shared memory. No global
// memory access is left inside an iteration apart from the command record (prefetched one iteration ahead) and thread 0's
// status / maximum stores. The per-voxel functions are the kernels' own (slav_gradient_at, slav_filter_axis_at,
// slav_resample_at): they address voxel idx of a field as base[component * N + idx], so each field gets a VIRTUAL base
// pointer into shared memory (slot of voxel 0 of the field, were the whole field there) and component 1 of a vector field
// sits N floats after component 0 (the four vector fields' component-0 tiles must fit in N floats: slav_strips_shape).
constexpr int STRIP_FIELDS = 7, STRIP_VECTOR_FIELDS = 4;
struct SlavStripFields {
	float* base[STRIP_FIELDS];  // global buffers: the vector fields (planes) first, then the scalar ones; the last is read-only
};

__device__ __forceinline__ float block_max(float value, float* warp_max) {
	if (!(value >= 0.0f)) value = 0.0f;  // NaN -> ignored, like block_atomic_max
#pragma unroll
	for (int offset = 16; offset > 0; offset >>= 1) value = fmaxf(value, __shfl_xor_sync(0xffffffffu, value, offset));
	if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = value;
	__syncthreads();
	if (threadIdx.x < 32) {
		value = threadIdx.x < ((blockDim.x + 31) >> 5) ? warp_max[threadIdx.x] : 0.0f;
#pragma unroll
		for (int offset = 16; offset > 0; offset >>= 1) value = fmaxf(value, __shfl_xor_sync(0xffffffffu, value, offset));
	}
	return value;  // valid in warp 0
}

// the pointer fields of a command record that name optimizer buffers: byte offset and the filter pass the field belongs to
// (-1: always used)
constexpr int COMMAND_POINTERS = 20;
__device__ __forceinline__ int command_pointer_pass(int t) {
	return t >= 5 && t < 14 ? (t - 5) / 3 : -1;
}
__device__ __forceinline__ size_t command_pointer_offset(int t) {
	switch (t) {
	case 0: return offsetof(SlavIterationCommand, gradient) + offsetof(SlavGradientArgs, live);
	case 1: return offsetof(SlavIterationCommand, gradient) + offsetof(SlavGradientArgs, canonical);
	case 2: return offsetof(SlavIterationCommand, gradient) + offsetof(SlavGradientArgs, warp);
	case 3: return offsetof(SlavIterationCommand, gradient) + offsetof(SlavGradientArgs, stale);
	case 4: return offsetof(SlavIterationCommand, gradient) + offsetof(SlavGradientArgs, out);
	case 5: case 8: case 11:
		return offsetof(SlavIterationCommand, pass) + (size_t) ((t - 5) / 3) * sizeof(SlavFilterArgs) + offsetof(SlavFilterArgs, in);
	case 6: case 9: case 12:
		return offsetof(SlavIterationCommand, pass) + (size_t) ((t - 5) / 3) * sizeof(SlavFilterArgs) + offsetof(SlavFilterArgs, out);
	case 7: case 10: case 13:
		return offsetof(SlavIterationCommand, pass) + (size_t) ((t - 5) / 3) * sizeof(SlavFilterArgs) + offsetof(SlavFilterArgs, original);
	case 14: return offsetof(SlavIterationCommand, resample) + offsetof(SlavResampleArgs, live);
	case 15: return offsetof(SlavIterationCommand, resample) + offsetof(SlavResampleArgs, canonical);
	case 16: return offsetof(SlavIterationCommand, resample) + offsetof(SlavResampleArgs, update);
	case 17: return offsetof(SlavIterationCommand, resample) + offsetof(SlavResampleArgs, gradient_field);
	case 18: return offsetof(SlavIterationCommand, resample) + offsetof(SlavResampleArgs, warp);
	default: return offsetof(SlavIterationCommand, resample) + offsetof(SlavResampleArgs, new_live);
	}
}

__global__ void __launch_bounds__(CLUSTER_THREADS, 1) k_slav_strips(const SlavIterationCommand* commands, int count, SlavParams p,
		int H, int W, int rows_per, int halo, SlavStripFields fields, unsigned* max_sq_bits, int* status, int first_iteration,
		int max_iterations, unsigned long long* phase_clock) {
	constexpr int D = 2;
	constexpr int WORDS = (int) (sizeof(SlavIterationCommand) / 4), PREFETCH = (WORDS + 127) / 128;  // >= 128 threads per block
	extern __shared__ __align__(16) float strip_tiles[];
	__shared__ SlavIterationCommand command;
	__shared__ float* virtual_base[STRIP_FIELDS];
	__shared__ float block_maxima[2][CLUSTER_BLOCKS], warp_max[32];  // one set of slots per iteration parity
	cg::cluster_group cluster = cg::this_cluster();
	const int rank = (int) cluster.block_rank(), blocks = (int) cluster.num_blocks();
	const int N = H * W, tile = (rows_per + 2 * halo) * W;
	const int r0 = rank * rows_per, r1 = min(r0 + rows_per, H);
	const int row_lo = max(r0 - halo, 0), row_hi = min(r1 + halo, H);
	if (threadIdx.x == 0) {
		slav_strip.enabled = 1;
		slav_strip.rank = rank;
		slav_strip.blocks = blocks;
		slav_strip.rows_per = rows_per;
		slav_strip.row_lo = row_lo;
		slav_strip.row_hi = row_hi;
		slav_strip.W = W;
		for (int f = 0; f < STRIP_FIELDS; f++) {
			const int offset = f < STRIP_VECTOR_FIELDS ? f * tile : N + f * tile;
			virtual_base[f] = strip_tiles + offset - (r0 - halo) * W;
		}
	}
	__syncthreads();
	// the block's rows (and halo rows) of every field
	for (int f = 0; f < STRIP_FIELDS; f++)
		for (int c = 0; c < (f < STRIP_VECTOR_FIELDS ? D : 1); c++)
			for (int idx = row_lo * W + threadIdx.x; idx < row_hi * W; idx += blockDim.x)
				virtual_base[f][c * N + idx] = fields.base[f][c * N + idx];
	// a command record names the global buffers: translated to the virtual bases when it is staged
	int prefetched[PREFETCH];
	auto prefetch = [&](int j) {
#pragma unroll
		for (int k = 0; k < PREFETCH; k++) {
			const int w = threadIdx.x + k * 128;
			if (threadIdx.x < 128 && w < WORDS && j < count) prefetched[k] = reinterpret_cast<const int*>(commands + j)[w];
		}
	};
	// thread t < COMMAND_POINTERS translates pointer field t of the staged record (the fields of passes that do not run are
	// left alone: they hold whatever the host's record held)
	auto stage_command = [&]() {
		// (the cluster barrier at the end of the previous iteration stands between that iteration's readers of `command`
		// and these stores)
#pragma unroll
		for (int k = 0; k < PREFETCH; k++) {
			const int w = threadIdx.x + k * 128;
			if (threadIdx.x < 128 && w < WORDS) reinterpret_cast<int*>(&command)[w] = prefetched[k];
		}
		__syncthreads();
		if (threadIdx.x < COMMAND_POINTERS && command_pointer_pass(threadIdx.x) < command.passes) {
			float** field = reinterpret_cast<float**>(reinterpret_cast<unsigned char*>(&command) + command_pointer_offset(threadIdx.x));
			const float* pointer = *field;
			if (pointer != nullptr) {
				float* translated = nullptr;
				for (int f = 0; f < STRIP_FIELDS; f++)
					if (pointer == fields.base[f]) translated = virtual_base[f];
				if (translated == nullptr) __trap();  // a buffer the host did not announce
				*field = translated;
			}
		}
		__syncthreads();
	};
	// the rows the neighbours keep as halo: stored into their shared memory (slot = this block's slot shifted by a strip)
	// (edges: bit 0 = the upper neighbour keeps the voxel's row as halo, bit 1 = the lower one)
	auto push = [&](float* base, int components, int idx, int edges) {
		const bool up = edges & 1, down = edges & 2;
		if (edges == 0) return;
		for (int c = 0; c < components; c++) {
			float* mine = base + c * N + idx;
			const float value = *mine;
			if (up) *cluster.map_shared_rank(mine + rows_per * W, rank - 1) = value;
			if (down) *cluster.map_shared_rank(mine - rows_per * W, rank + 1) = value;
		}
	};
	// the tap loop unrolled for the usual radii (the launch's filter radius = halo, when there is a filter)
	auto filter_pass = [&](const SlavFilterArgs& fa, int idx, const int (&pos)[3]) {
		if (fa.radius == 3) slav_filter_axis_at<D, 3>(fa, idx, pos);
		else if (fa.radius == 1) slav_filter_axis_at<D, 1>(fa, idx, pos);
		else if (fa.radius == 2) slav_filter_axis_at<D, 2>(fa, idx, pos);
		else slav_filter_axis_at<D>(fa, idx, pos);
	};
	const bool stamping = phase_clock != nullptr && rank == 0 && threadIdx.x == 0;
	auto stamp = [&](int j, int slot) {
		if (stamping && j < PHASE_CLOCK_ITERATIONS) {
			unsigned long long now;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
			phase_clock[j * 8 + slot] = now;
		}
	};
	const int first = r0 * W + threadIdx.x, last = r1 * W;
	// coordinates of a thread's voxels: the first one's are computed once per launch (blocks usually have a thread per voxel)
	const int first_pos[3] = { first / W, first % W, 0 };
	auto row_edges = [&](int row) { return (rank > 0 && row - r0 < halo ? 1 : 0) | (rank + 1 < blocks && r1 - 1 - row < halo ? 2 : 0); };
	const int first_edges = row_edges(first_pos[0]);
	auto coordinates = [&](int idx, int (&pos)[3]) {
		pos[0] = idx == first ? first_pos[0] : idx / W;
		pos[1] = idx == first ? first_pos[1] : idx - pos[0] * W;
		pos[2] = 0;
		return idx == first ? first_edges : row_edges(pos[0]);
	};
	bool finished = *reinterpret_cast<volatile int*>(status + first_iteration) != 0;
	prefetch(0);
	cluster.sync();  // every block has loaded its tiles: halo stores may arrive from now on
	int j = 0;
	for (; j < count; j++) {
		const int it = first_iteration + j;
		if (finished) break;
		stage_command();
		prefetch(j + 1);
		stamp(j, 0);
		const int passes = command.passes;
		// Only a pass along axis 0 (rows) reads rows of other blocks: its input is pushed into the neighbours' halo and a
		// cluster barrier stands in front of it. A pass along axis 1 reads its own rows: nothing is pushed for it and a
		// block-wide barrier is enough (the blocks drift apart there and meet again at the end of the iteration).
		const bool first_pass_crosses = passes > 0 && command.pass[0].axis == 0;
		for (int idx = first; idx < last; idx += blockDim.x) {
			int pos[3];
			const int edges = coordinates(idx, pos);
			slav_gradient_at<D>(command.gradient, idx, pos);
			if (first_pass_crosses) push(command.gradient.out, D, idx, edges);
		}
		stamp(j, 1);
		for (int pass = 0; pass + 1 < passes; pass++) {
			if (command.pass[pass].axis == 0) cluster.sync();
			else __syncthreads();
			if (pass == 0) stamp(j, 2);
			const bool next_crosses = command.pass[pass + 1].axis == 0;
			for (int idx = first; idx < last; idx += blockDim.x) {
				int pos[3];
				const int edges = coordinates(idx, pos);
				filter_pass(command.pass[pass], idx, pos);
				if (next_crosses) push(command.pass[pass].out, D, idx, edges);
			}
			if (pass == 0) stamp(j, 3);
		}
		// without a filter the cluster barrier stands between the neighbours' smoothing terms (gradient phase), which read
		// this block's warp vectors in their halo, and the re-warp that overwrites them; with a filter an earlier one does
		if (passes == 0 || command.pass[passes - 1].axis == 0) cluster.sync();
		else __syncthreads();
		stamp(j, 4);
		float sq_report = 0.0f;
		for (int idx = first; idx < last; idx += blockDim.x) {
			int pos[3];
			const int edges = coordinates(idx, pos);
			if (passes > 0) filter_pass(command.pass[passes - 1], idx, pos);
			slav_resample_at<D>(command.resample, idx, sq_report, pos);
			push(command.resample.new_live, 1, idx, edges);
			push(command.resample.warp, D, idx, edges);
		}
		stamp(j, 5);
		// maximum warp length: every block's maximum into every block's slot array
		const float mine = block_max(sq_report, warp_max);
		if (threadIdx.x < blocks) *cluster.map_shared_rank(&block_maxima[j & 1][rank], threadIdx.x) = mine;
		cluster.sync();
		stamp(j, 6);
		float max_sq = 0.0f;
		for (int k = 0; k < blocks; k++) max_sq = fmaxf(max_sq, block_maxima[j & 1][k]);
		finished = slav_finished(p, it + 1, max_iterations, sqrtf(max_sq));
		if (rank == 0 && threadIdx.x == 0) {
			max_sq_bits[it] = __float_as_uint(max_sq);
			status[it + 1] = finished ? 1 : 0;
		}
		stamp(j, 7);
	}
	if (finished && rank == 0 && threadIdx.x == 0)
		for (int k = j; k < count; k++) status[first_iteration + k + 1] = 1;  // sticky, like k_slav_decide
	// the block's rows of every buffer back to global memory (the host reads the results there; the next launch reloads them)
	for (int f = 0; f + 1 < STRIP_FIELDS; f++)
		for (int c = 0; c < (f < STRIP_VECTOR_FIELDS ? D : 1); c++)
			for (int idx = first; idx < last; idx += blockDim.x) fields.base[f][c * N + idx] = virtual_base[f][c * N + idx];
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

// strips of the distributed-shared-memory kernel for this field, if it takes it: every block at least `halo` rows (its halo
// comes from the direct neighbours only), the vector fields' component-0 tiles inside the first N floats, the cluster
// schedulable with that much shared memory. LSF_SLAV_CLUSTER=1 keeps the fields in global memory (A/B tests).
struct StripShape {
	unsigned blocks, threads;
	int rows_per, halo;
	size_t shared_bytes;
};
bool slav_strips_shape(const SlavOptimizerBuffers& b, StripShape* shape) {
	const char* env = getenv("LSF_SLAV_CLUSTER");
	if (env && (env[0] == '0' || env[0] == '1')) return false;
	const int most = cluster_blocks();
	if (most <= 0 || b.H < 2 || b.W < 1) return false;
	const int halo = std::max(b.radius, 1);
	const int rows_per = std::max((b.H + most - 1) / most, halo);
	const int blocks = (b.H + rows_per - 1) / rows_per;
	const long long N = (long long) b.H * b.W, tile = (long long) (rows_per + 2 * halo) * b.W;
	if (blocks < 1 || blocks > most || STRIP_VECTOR_FIELDS * tile > N) return false;
	const size_t bytes = (size_t) (N + STRIP_FIELDS * tile) * sizeof(float);
	if (bytes > 200u * 1024u) return false;
	unsigned threads = 128;
	while (threads < (unsigned) CLUSTER_THREADS && (long long) threads < (long long) rows_per * b.W) threads *= 2;
	// schedulable? (asked once per shape)
	static std::mutex mutex;
	static std::map<std::tuple<int, unsigned, size_t>, bool> known;
	std::lock_guard<std::mutex> lock(mutex);
	int device = 0;
	cudaGetDevice(&device);  // function attributes are per device
	const auto key = std::make_tuple(blocks * 64 + device, threads, bytes);
	auto found = known.find(key);
	if (found == known.end()) {
		bool ok = cudaFuncSetAttribute(k_slav_strips, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess
				&& (blocks <= 8 || cudaFuncSetAttribute(k_slav_strips, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess);
		cudaLaunchConfig_t config = {};
		config.gridDim = dim3(blocks);
		config.blockDim = dim3(threads);
		config.dynamicSmemBytes = bytes;
		cudaLaunchAttribute attribute;
		attribute.id = cudaLaunchAttributeClusterDimension;
		attribute.val.clusterDim.x = blocks;
		attribute.val.clusterDim.y = attribute.val.clusterDim.z = 1;
		config.attrs = &attribute;
		config.numAttrs = 1;
		int clusters = 0;
		ok = ok && cudaOccupancyMaxActiveClusters(&clusters, k_slav_strips, &config) == cudaSuccess && clusters >= 1;
		cudaGetLastError();
		found = known.emplace(key, ok).first;
	}
	if (!found->second) return false;
	shape->blocks = (unsigned) blocks;
	shape->threads = threads;
	shape->rows_per = rows_per;
	shape->halo = halo;
	shape->shared_bytes = bytes;
	return true;
}

}  // namespace

long long slav_persistent_capacity() {
	const int blocks = resident_blocks();
	// one voxel per thread up to a full wave; beyond that the ordinary kernels are no longer launch-bound
	return blocks > 0 ? (long long) blocks * PERSISTENT_THREADS : 0;
}

int launch_slav_persistent2d(const SlavIterationCommand* commands_dev, int count, const SlavParams& p, long long N,
		unsigned* max_sq_bits, int* status, int first_iteration, int max_iterations, cudaStream_t stream,
		const SlavOptimizerBuffers* buffers) {
	// LSF_TRACE=2: phase clock of the first chunk's first iterations, printed when the launch has finished (debugging aid:
	// synchronises the stream)
	static const bool clocked = []() {
		const char* e = getenv("LSF_TRACE");
		return e && e[0] == '2';
	}();
	unsigned long long* phase_clock = nullptr;
	if (clocked && first_iteration == 0) {
		LSF_CUDA(cudaMalloc(&phase_clock, PHASE_CLOCK_ITERATIONS * 8 * sizeof(unsigned long long)));
		LSF_CUDA(cudaMemsetAsync(phase_clock, 0, PHASE_CLOCK_ITERATIONS * 8 * sizeof(unsigned long long), stream));
	}
	auto report = [&]() {
		if (phase_clock == nullptr) return;
		unsigned long long host[PHASE_CLOCK_ITERATIONS * 8];
		cudaStreamSynchronize(stream);
		cudaMemcpy(host, phase_clock, sizeof(host), cudaMemcpyDeviceToHost);
		cudaFree(phase_clock);
		static const char* names[8] = { "command", "gradient", "barrier", "pass 0", "barrier", "pass 1 + re-warp", "max + barrier", "decide" };
		fprintf(stderr, "[lsf_b200 phase clock] largest cluster %d blocks, %lld voxels\n", cluster_blocks(), N);
		for (int j = 1; j < PHASE_CLOCK_ITERATIONS && j < count; j++) {
			fprintf(stderr, "[lsf_b200 phase clock] iteration %d:", j);
			unsigned long long previous = host[(j - 1) * 8 + 7];
			for (int slot = 0; slot < 8; slot++) {
				if (host[j * 8 + slot] == 0) continue;
				fprintf(stderr, " %s %llu ns |", names[slot], host[j * 8 + slot] - previous);
				previous = host[j * 8 + slot];
			}
			fprintf(stderr, " total %llu ns\n", host[j * 8 + 7] - host[(j - 1) * 8 + 7]);
		}
	};
	LSF_REQUIRE(N > 0 && N <= slav_persistent_capacity(), "field of %lld voxels does not fit the single-launch path", N);
	int n = (int) N;
	SlavParams params = p;
	StripShape shape;
	if (buffers != nullptr && slav_strips_shape(*buffers, &shape)) {
		SlavStripFields fields;
		for (int f = 0; f < STRIP_VECTOR_FIELDS; f++) fields.base[f] = buffers->vector_fields[f];
		for (int f = STRIP_VECTOR_FIELDS; f < STRIP_FIELDS; f++) fields.base[f] = buffers->scalar_fields[f - STRIP_VECTOR_FIELDS];
		cudaLaunchConfig_t config = {};
		config.gridDim = dim3(counted(shape.blocks));
		config.blockDim = dim3(shape.threads);
		config.dynamicSmemBytes = shape.shared_bytes;
		config.stream = stream;
		cudaLaunchAttribute attribute;
		attribute.id = cudaLaunchAttributeClusterDimension;
		attribute.val.clusterDim.x = shape.blocks;
		attribute.val.clusterDim.y = attribute.val.clusterDim.z = 1;
		config.attrs = &attribute;
		config.numAttrs = 1;
		LSF_CUDA(cudaLaunchKernelEx(&config, k_slav_strips, commands_dev, count, params, buffers->H, buffers->W, shape.rows_per,
				shape.halo, fields, max_sq_bits, status, first_iteration, max_iterations, phase_clock));
		report();
		return LSF_OK;
	}
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
				max_iterations, phase_clock));
		report();
		return LSF_OK;
	}
	const unsigned blocks = (unsigned) std::min<long long>(div_up(N, PERSISTENT_THREADS), resident_blocks());
	void* arguments[] = { (void*) &commands_dev, (void*) &count, (void*) &params, (void*) &n, (void*) &max_sq_bits, (void*) &status,
			(void*) &first_iteration, (void*) &max_iterations, (void*) &phase_clock };
	LSF_CUDA(cudaLaunchCooperativeKernel((const void*) k_slav_persistent<2, false>, dim3(counted(blocks)), dim3(PERSISTENT_THREADS), arguments,
			0, stream));
	report();
	return LSF_OK;
}

}  // namespace lsf
