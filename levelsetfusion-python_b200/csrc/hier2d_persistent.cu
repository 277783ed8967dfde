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

#include <cuda_runtime.h>

namespace lsf {
namespace {
// the strip of the padded live pack a block of the distributed-shared-memory level kernel (k_hier_level2d_strips, below)
// holds: padded rows [pack_lo, pack_hi) are in this block's shared memory; padded row p belongs to block (p - 2) / rows_per
struct HierStrip {
	int enabled;  // 0: the fields live in global memory (k_hier_level2d)
	int rank, blocks, rows_per, pack_lo, pack_hi, PW;
};
__shared__ HierStrip hier_strip;

// a tap of the gather: padded rows outside the block's strip + halo are read from the shared memory of the block that owns
// them (all blocks lay their tiles out the same way: the owner's slot of an element is this block's slot shifted by the
// distance of the two strips)
__device__ __forceinline__ float4 hier_strip_pack_tap(const float4* pack, int padded_row, long long index) {
	if (!hier_strip.enabled || (padded_row >= hier_strip.pack_lo && padded_row < hier_strip.pack_hi)) return pack[index];
	const int owner = min(max((padded_row - 2) / hier_strip.rows_per, 0), hier_strip.blocks - 1);
	const float4* slot = pack + index + (long long) (hier_strip.rank - owner) * hier_strip.rows_per * hier_strip.PW;
	return *cooperative_groups::this_cluster().map_shared_rank(slot, owner);
}
}  // namespace
}  // namespace lsf
#define HIER_PACK_TAP(pack, padded_row, index) lsf::hier_strip_pack_tap(pack, padded_row, index)
#include "kernels2d.cuh"

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>

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
	if (threadIdx.x == 0) hier_strip.enabled = 0;  // the fields are in global memory
	__syncthreads();
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


// ---------------------------------------------------------------------------------------------- fields in distributed shared memory
// The small pyramid levels are bound by latency: an iteration of k_hier_level2d costs 2.2 us per phase whatever the level's
// size (one barrier + one chain of dependent L2 loads; tools/overhead2d.py). k_hier_level2d_strips keeps the level in the
// shared memory of ONE cluster: block k owns the rows [k * rows_per, (k + 1) * rows_per) of the warp field, the canonical
// field and the two gradient fields (these with `halo` rows either side: filter radius / 1 for the Laplacian) and the
// matching rows of the padded live pack (PACK_HALO rows either side). A phase reads this block's shared memory only, writes
// its rows and stores the rows its neighbours keep as halo into THEIR shared memory; cluster barriers separate the phases;
// gather taps outside the block's pack rows are read from their owner's shared memory (hier_strip_pack_tap); the maximum of
// an iteration travels through a slot per block in every block's shared memory. Per-pixel functions = the kernels' own
// (hier_gradient2d_at, convolve_axis2d_at): every field gets a virtual base pointer into shared memory, the component
// stride of the two-plane fields (Grid2::N in the argument blocks) is the tile size. Only the warp field is written back.
constexpr int PACK_HALO = 4;

__device__ __forceinline__ float strip_block_max(float value, float* warp_max) {
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

template<bool TIKHONOV, int R>
__global__ void __launch_bounds__(CLUSTER_THREADS, 1) k_hier_level2d_strips(HierIterArgs2 a, ConvArgs2 c, int use_kernel,
		float* g_post_global, float* scratch_global, int rows_per, int halo, int first_iteration, int count) {
	extern __shared__ __align__(16) unsigned char strip_memory[];
	// slots of the blocks' maxima, one set per iteration parity: without a filter nothing but this exchange's own barrier
	// separates two iterations, and a block that is ahead stores its next maximum while a slower block still reads the last
	__shared__ float block_maxima[3][CLUSTER_BLOCKS], warp_max[2][32];  // (three / two sets: the data-term-only loop below)
	cg::cluster_group cluster = cg::this_cluster();
	const int rank = (int) cluster.block_rank(), blocks = (int) cluster.num_blocks();
	const Grid2 g = a.g;
	const int H = g.H, W = g.W, PW = g.PW();
	const int r0 = rank * rows_per, r1 = min(r0 + rows_per, H);
	const int tile = (rows_per + 2 * halo) * W;                      // floats per plane
	const int pack_rows = rows_per + 2 * PACK_HALO + 1;              // padded rows r0 + 2 - PACK_HALO ... (unclipped)
	const int pack_first = r0 + 2 - PACK_HALO;
	const int pack_lo = max(pack_first, 0), pack_hi = min(pack_first + pack_rows, H + 4);
	if (threadIdx.x == 0) {
		hier_strip.enabled = 1;
		hier_strip.rank = rank;
		hier_strip.blocks = blocks;
		hier_strip.rows_per = rows_per;
		hier_strip.pack_lo = pack_lo;
		hier_strip.pack_hi = pack_hi;
		hier_strip.PW = PW;
	}
	// tiles: pack | warp (2 planes) | g_post (2) | scratch (2) | canonical | between (2); virtual base = slot of element 0 of the field
	float4* pack_tile = reinterpret_cast<float4*>(strip_memory);
	float* planes = reinterpret_cast<float*>(pack_tile + (size_t) pack_rows * PW);
	const long long shift = (long long) (r0 - halo) * W;
	float4* pack = pack_tile - (long long) pack_first * PW;
	float* warp = planes - shift;
	float* g_post = planes + 2 * tile - shift;
	float* scratch = planes + 4 * tile - shift;
	float* canonical = planes + 6 * tile - shift;
	float* between = planes + 7 * tile - shift;  // output of the rows pass (two planes, never leaves the block)
	const long long N = g.N;
	const int row_lo = max(r0 - halo, 0), row_hi = min(r1 + halo, H);
	for (long long i = (long long) pack_lo * PW + threadIdx.x; i < (long long) pack_hi * PW; i += blockDim.x) pack[i] = a.pack[i];
	for (long long idx = (long long) row_lo * W + threadIdx.x; idx < (long long) row_hi * W; idx += blockDim.x) {
		for (int k = 0; k < 2; k++) {
			warp[k * tile + idx] = a.warp[k * N + idx];
			g_post[k * tile + idx] = g_post_global[k * N + idx];
			scratch[k * tile + idx] = scratch_global[k * N + idx];
		}
		canonical[idx] = a.canonical[idx];
	}
	const float* warp_global_in = a.warp;
	float* warp_global = a.warp_out;
	unsigned* max_sq_bits = a.max_sq_bits;
	// the argument blocks now name the tiles; the component stride of the two-plane fields is the tile size
	a.pack = pack;
	a.canonical = canonical;
	a.warp = warp;
	a.warp_out = warp;
	a.g.N = tile;
	c.warp = warp;
	c.g.N = tile;
	(void) warp_global_in;
	// rows the neighbours keep as halo go into their shared memory (slot = this block's slot shifted by a strip)
	auto push = [&](float* base, long long idx, int edges) {
		if (edges == 0) return;
		for (int k = 0; k < 2; k++) {
			float* mine = base + k * tile + idx;
			const float value = *mine;
			if (edges & 1) *cluster.map_shared_rank(mine + rows_per * W, rank - 1) = value;
			if (edges & 2) *cluster.map_shared_rank(mine - rows_per * W, rank + 1) = value;
		}
	};
	auto row_edges = [&](int row) { return (rank > 0 && row - r0 < halo ? 1 : 0) | (rank + 1 < blocks && r1 - 1 - row < halo ? 2 : 0); };
	const long long first = (long long) r0 * W + threadIdx.x, last = (long long) r1 * W;
	// row, column and halo flags of a thread's pixels: the first one's are computed once per launch (blocks usually have a
	// thread per pixel)
	const int first_row = (int) (first / W), first_col = (int) (first - (long long) first_row * W), first_edges = row_edges(first_row);
	auto coordinates = [&](long long idx, int& row, int& col) {
		if (idx == first) {
			row = first_row;
			col = first_col;
			return first_edges;
		}
		row = (int) (idx / W);
		col = (int) (idx - (long long) row * W);
		return row_edges(row);
	};
	float previous_max_sq = 0.0f;
	if (first_iteration > 0) previous_max_sq = __uint_as_float(*reinterpret_cast<const volatile unsigned*>(max_sq_bits + first_iteration - 1));
	__syncthreads();
	cluster.sync();  // every block has loaded its tiles: halo stores may arrive from now on
	if (!use_kernel && !TIKHONOV) {
		// Data term only: a pixel's iteration reads and writes that pixel's warp vector alone, so nothing but the termination
		// test (optimizer.tpp:166-171: stop after the first iteration whose maximum is below the threshold) connects the
		// blocks. The test runs one iteration late: iteration `it` is computed while the maxima of iteration it - 1 travel
		// (split cluster barrier: arrive after storing the maximum, wait after the next iteration's work), into the OTHER of
		// two warp tiles (the unused gradient tile), so that an iteration computed past the stopping point is simply dropped.
		// Slot sets: a block that is ahead stores the maximum of it + 1 while a slower one still reads it - 1 -> three sets.
		float* tiles[2] = { warp, g_post };
		int current = 0;  // tiles[current] = the warp field before iteration `it`
		bool pending = false;
		a.g_prev = nullptr;
		a.g_out = nullptr;
		int it = first_iteration;
		if (it > 0 && sqrtf(previous_max_sq) < a.threshold) it = first_iteration + count;  // nothing left to do
		for (; it < first_iteration + count; it++) {
			a.warp = tiles[current];
			a.warp_out = tiles[current ^ 1];
			float sq = 0.0f;
			for (long long idx = first; idx < last; idx += blockDim.x) {
				int row, col;
				coordinates(idx, row, col);
				float mine = 0.0f;
				hier_gradient2d_at<false, true>(a, row, col, idx, mine);
				sq = fmaxf(sq, mine);
			}
			// (no block-wide barrier follows the reduction in this loop: its scratch alternates, a warp cannot be two iterations ahead)
			const float mine = strip_block_max(sq, warp_max[it & 1]);
			if (threadIdx.x < blocks) *cluster.map_shared_rank(&block_maxima[it % 3][rank], threadIdx.x) = mine;
			current ^= 1;
			if (pending) {
				cluster.barrier_wait();  // the maxima of iteration it - 1 have arrived
				float before = 0.0f;
				for (int k = 0; k < blocks; k++) before = fmaxf(before, block_maxima[(it - 1) % 3][k]);
				if (rank == 0 && threadIdx.x == 0) max_sq_bits[it - 1] = __float_as_uint(before);
				if (sqrtf(before) < a.threshold) {
					current ^= 1;  // iteration `it` ran past the stopping point: dropped
					pending = false;
					break;
				}
			}
			cluster.barrier_arrive();
			pending = true;
		}
		if (pending) {
			cluster.barrier_wait();
			float before = 0.0f;
			for (int k = 0; k < blocks; k++) before = fmaxf(before, block_maxima[(it - 1) % 3][k]);
			if (rank == 0 && threadIdx.x == 0) max_sq_bits[it - 1] = __float_as_uint(before);
		}
		cluster.sync();  // no block leaves while maxima of a dropped iteration may still be on their way into its slots
		for (long long idx = first; idx < last; idx += blockDim.x)
			for (int k = 0; k < 2; k++) warp_global[k * N + idx] = tiles[current][k * tile + idx];
		return;
	}
	for (int it = first_iteration; it < first_iteration + count; it++) {
		if (it > 0 && sqrtf(previous_max_sq) < a.threshold) break;  // level_converged()
		a.g_prev = g_post;
		float sq = 0.0f;
		if (!use_kernel) {
			a.g_out = TIKHONOV ? scratch : nullptr;
			for (long long idx = first; idx < last; idx += blockDim.x) {
				int row, col;
				const int edges = coordinates(idx, row, col);
				float mine = 0.0f;
				hier_gradient2d_at<TIKHONOV, true>(a, row, col, idx, mine);
				sq = fmaxf(sq, mine);
				if (TIKHONOV) push(scratch, idx, edges);
			}
			if (TIKHONOV) {
				float* t = g_post;
				g_post = scratch;
				scratch = t;
			}
		} else {
			a.g_out = scratch;
			for (long long idx = first; idx < last; idx += blockDim.x) {
				int row, col;
				const int edges = coordinates(idx, row, col);
				float unused = 0.0f;
				hier_gradient2d_at<TIKHONOV, false>(a, row, col, idx, unused);
				push(scratch, idx, edges);
			}
			// Only the rows pass reads rows of other blocks (its input was pushed above, a cluster barrier in front of it);
			// the columns pass reads its own rows of the rows pass's output, which stays in the block (block barrier), and
			// writes the filtered gradient back into g_post -- every reader of the old one (stage 1 above, any block) is
			// behind the cluster barrier -- so the two fields keep their roles; its rows go to the neighbours' halo for
			// the next iteration's Laplacian.
			cluster.sync();
			c.in = scratch;
			c.out = between;
			for (long long idx = first; idx < last; idx += blockDim.x) {
				int row, col;
				coordinates(idx, row, col);
				float unused = 0.0f;
				convolve_axis2d_at<0, false, R>(c, row, col, idx, unused);
			}
			__syncthreads();
			c.in = between;
			c.out = g_post;
			for (long long idx = first; idx < last; idx += blockDim.x) {
				int row, col;
				const int edges = coordinates(idx, row, col);
				float mine = 0.0f;
				convolve_axis2d_at<1, true, R>(c, row, col, idx, mine);
				sq = fmaxf(sq, mine);
				if (TIKHONOV) push(g_post, idx, edges);
			}
		}
		// the iteration's maximum: every block's maximum into every block's slot array
		const float mine = strip_block_max(sq, warp_max[0]);
		if (threadIdx.x < blocks) *cluster.map_shared_rank(&block_maxima[it & 1][rank], threadIdx.x) = mine;
		cluster.sync();
		previous_max_sq = 0.0f;
		for (int k = 0; k < blocks; k++) previous_max_sq = fmaxf(previous_max_sq, block_maxima[it & 1][k]);
		if (rank == 0 && threadIdx.x == 0) max_sq_bits[it] = __float_as_uint(previous_max_sq);
	}
	for (long long idx = first; idx < last; idx += blockDim.x)
		for (int k = 0; k < 2; k++) warp_global[k * N + idx] = warp[k * tile + idx];
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

// strips of the distributed-shared-memory level kernel for this level, if it takes it: every block at least `halo` rows (its
// halo comes from the direct neighbours only), the tiles within the shared memory of a block, the cluster schedulable.
// LSF_HIER2D_CLUSTER=1 keeps the fields in global memory (A/B tests).
struct StripShape {
	unsigned blocks, threads;
	int rows_per, halo;
	size_t shared_bytes;
};
// the tap loops are unrolled for the usual filter radii (0 = any / no filter: runtime loop)
typedef void (*StripsKernel)(HierIterArgs2, ConvArgs2, int, float*, float*, int, int, int, int);
StripsKernel strips_kernel(bool tikhonov, int radius) {
	switch (radius) {
	case 1: return tikhonov ? k_hier_level2d_strips<true, 1> : k_hier_level2d_strips<false, 1>;
	case 2: return tikhonov ? k_hier_level2d_strips<true, 2> : k_hier_level2d_strips<false, 2>;
	case 3: return tikhonov ? k_hier_level2d_strips<true, 3> : k_hier_level2d_strips<false, 3>;
	default: return tikhonov ? k_hier_level2d_strips<true, 0> : k_hier_level2d_strips<false, 0>;
	}
}
bool hier2d_strips_shape(const Grid2& g, bool tikhonov, int radius, StripShape* shape) {
	const char* env = getenv("LSF_HIER2D_CLUSTER");
	if (env && (env[0] == '0' || env[0] == '1')) return false;
	const int most = cluster_blocks();
	if (most <= 0 || g.H < 1 || g.W < 1) return false;
	const int halo = std::max(radius, 1);
	const int rows_per = std::max((g.H + most - 1) / most, halo);
	const int blocks = (g.H + rows_per - 1) / rows_per;
	if (blocks < 1 || blocks > most) return false;
	const size_t tile = (size_t) (rows_per + 2 * halo) * g.W;
	const size_t bytes = (size_t) (rows_per + 2 * PACK_HALO + 1) * g.PW() * sizeof(float4) + 9 * tile * sizeof(float);
	if (bytes > 200u * 1024u) return false;
	unsigned threads = 128;
	while (threads < (unsigned) CLUSTER_THREADS && (long long) threads < (long long) rows_per * g.W) threads *= 2;
	// schedulable? (asked once per shape and kernel)
	static std::mutex mutex;
	static std::map<std::tuple<int, unsigned, size_t, int>, bool> known;
	std::lock_guard<std::mutex> lock(mutex);
	int device = 0;
	cudaGetDevice(&device);  // function attributes are per device
	const auto key = std::make_tuple(blocks, threads, bytes, device * 16 + (tikhonov ? 8 : 0) + std::min(radius, 4));
	auto found = known.find(key);
	if (found == known.end()) {
		const void* kernel = reinterpret_cast<const void*>(strips_kernel(tikhonov, radius));
		bool ok = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess
				&& (blocks <= 8 || cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess);
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
		ok = ok && cudaOccupancyMaxActiveClusters(&clusters, kernel, &config) == cudaSuccess && clusters >= 1;
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
	StripShape shape;
	// (a level is one launch: the strips kernel does not write its gradient fields back for a later chunk)
	if (first_iteration == 0 && hier2d_strips_shape(gradient.g, tikhonov, use_kernel ? filter.taps.radius : 0, &shape)) {
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
		LSF_CUDA(cudaLaunchKernelEx(&config, strips_kernel(tikhonov, use_kernel ? filter.taps.radius : 0), a, c, kernel_flag, g_post, scratch,
				shape.rows_per, shape.halo, first_iteration, count));
		return LSF_OK;
	}
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
