// hier3d.cu -- 3D hierarchical optimizer driver (host side of the C-ABI) on top of kernels3d.cuh.
//
// Restates the control flow of the reference's Optimizer<Tensor3f,Tensor3v3f>
// (cpp/src/nonrigid_optimization/hierarchical/optimizer.tpp:83-212, pyramid.tpp:51-74) around fused
// sm_100a kernels. The level loop never synchronises with the host inside an iteration: the
// termination test (optimizer.tpp:166-171) is evaluated on the device at the head of every kernel from
// the previous iteration's max ||g||^2 slot, and the host polls the slots once per chunk of iterations.
#include "kernels3d_fused.cuh"
#include "kernels3d_split.cuh"
#include "kernels3d_pair.cuh"
#include "kernels3d_ymarch3.cuh"
#include "slavcheva.cuh"  // statistics_on_device
#include "hier_telemetry.cuh"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <cstdint>
#include <algorithm>
#include <atomic>

namespace lsf {

namespace {

// which kernel family the last enqueued 3D iteration used (lsf_debug_last_path; the parity tests assert that they
// exercised the kernels bench.py measures)
std::atomic<int> g_last_path { 0 };

constexpr int POLL_CHUNK = 16;  // iterations enqueued between two host polls of the convergence slots

// Per-thread polling resources of optimize(): a pinned staging buffer for the convergence slots (a pageable destination
// would make cudaMemcpyAsync block the host) and two events, one per chunk in flight.
struct PollState {
	unsigned* host_bits = nullptr;
	size_t capacity = 0;
	cudaEvent_t events[2] = { nullptr, nullptr };
	int device = -1;
	int reserve(size_t count) {
		int current = 0;
		LSF_CUDA(cudaGetDevice(&current));
		if (events[0] != nullptr && current != device) {  // the calling thread moved to another GPU
			cudaEventDestroy(events[0]);
			cudaEventDestroy(events[1]);
			events[0] = events[1] = nullptr;
		}
		device = current;
		if (events[0] == nullptr) {
			LSF_CUDA(cudaEventCreateWithFlags(&events[0], cudaEventDisableTiming));
			LSF_CUDA(cudaEventCreateWithFlags(&events[1], cudaEventDisableTiming));
		}
		if (count > capacity) {
			if (host_bits) cudaFreeHost(host_bits);
			host_bits = nullptr;
			capacity = 0;
			LSF_CUDA(cudaMallocHost(&host_bits, count * sizeof(unsigned)));
			capacity = count;
		}
		return LSF_OK;
	}
};
PollState& poll_state() {
	static thread_local PollState state;
	return state;
}

struct Plan3 {
	bool tikhonov = false, use_kernel = false, linear = false;
	int level_count = 0;
	Taps taps;
	float rate = 0, threshold = 0, amplifier = 0, strength = 0;
	int max_iterations = 0;
	bool allow_fast_kernels = true;    // false: first-generation kernels only (kept for A/B parity tests)
	int lane_xv = 4;                   // planes per thread of the lane-contiguous stage 1 (LSF_LANE_XV overrides;
	                                   // measured best on B200: 2 with the Tikhonov term, 4 without)
	bool split_x = true;               // stage 1 + axis-0 pass | axis-1/2 passes + update (LSF_SPLIT_X=0: previous cut)
	int x_chunk_stage1 = 64, x_chunk_filter = 16;  // planes per block of the two split kernels (LSF_XCHUNK_A / _B)
	int stage1_variant = 2;            // LSF_STAGE1_VARIANT=1 selects the 4-voxel kernel (A/B)
	bool slab_fast = true;             // slab mode: TMA-fed stage 1 and marching filter kernels (LSF_SLAB_FAST=0: first generation)
	bool tma = true;                   // third generation: TMA-fed stage 1 + y-marching filter (LSF_TMA=0: second generation)
	int pair_tile_y = 0;               // fourth generation: two voxels per thread in stage 1, 64 x pair_tile_y tiles (LSF_PAIR_TY=0: third generation)
	int x_chunk_tma = 0, y_chunk_tma = 0;   // planes / rows per block of the two TMA-generation kernels (LSF_XCHUNK_T /
	                                        // LSF_YCHUNK_T); 0 = chosen per level by marching_chunk()
	Grid3 level_grid[LSF_MAX_LEVELS];  // [0] = coarsest
};

int make_plan(const lsf_hier_params* p, int X, int Y, int Z, Plan3* plan) {
	LSF_REQUIRE(p != nullptr, "params is NULL");
	LSF_REQUIRE(X > 0 && Y > 0 && Z > 0, "field dimensions must be positive, got %d x %d x %d", X, Y, Z);
	// reference pyramid.tpp:53-60
	LSF_REQUIRE(is_power_of_two(p->maximum_chunk_size),
			"The argument 'maximum_chunk_size' must be an integer power of 2, i.e. 4, 8, 16, etc.");
	const int power = (int) std::log2((double) p->maximum_chunk_size);
	const int max_level_count = (int) std::min( { std::log2((double) X), std::log2((double) Y), std::log2(
			(double) Z) }) + 1;
	LSF_REQUIRE(max_level_count > power, "Maximum chunk size too large for the field size.");
	plan->level_count = power + 1;
	LSF_REQUIRE(plan->level_count <= LSF_MAX_LEVELS, "too many pyramid levels (%d)", plan->level_count);
	plan->linear = p->resampling_strategy == LSF_RESAMPLING_LINEAR;
	LSF_REQUIRE(p->resampling_strategy == LSF_RESAMPLING_LINEAR
			|| p->resampling_strategy == LSF_RESAMPLING_NEAREST_AND_AVERAGE, "Unknown resampling strategy %d",
			p->resampling_strategy);
	const int divisor = 1 << (plan->level_count - 1);
	// the reference halves with integer division and later doubles again (resampling.tpp:388-392,107-108):
	// only sizes divisible by 2^(levels-1) survive the round trip
	LSF_REQUIRE(X % divisor == 0 && Y % divisor == 0 && Z % divisor == 0,
			"each dimension (%d x %d x %d) must be divisible by %d for a %d-level pyramid", X, Y, Z, divisor,
			plan->level_count);
	Grid3 g(X, Y, Z);
	for (int level = plan->level_count - 1; level >= 0; level--) {
		plan->level_grid[level] = g;
		if (level > 0 && plan->linear) {
			// reference resampling.tpp:547-549
			LSF_REQUIRE(g.X % 2 == 0 && g.Y % 2 == 0 && g.Z % 2 == 0 && g.X > 2 && g.Y > 2 && g.Z > 2,
					"Each dimension of the argument 'field' must be divisible by 2 and greater than 2.");
		}
		g = g.half();
	}
	// reference optimizer.tpp:65-66
	plan->tikhonov = p->tikhonov_term_enabled && p->tikhonov_strength > 0.0f;
	plan->use_kernel = p->gradient_kernel_enabled && p->kernel_size > 0 && p->kernel != nullptr;
	if (plan->use_kernel) LSF_TRY(make_taps(p->kernel, p->kernel_size, &plan->taps));
	plan->rate = p->rate;
	plan->threshold = p->maximum_warp_update_threshold;
	plan->amplifier = p->data_term_amplifier;
	plan->strength = p->tikhonov_strength;
	plan->max_iterations = p->maximum_iteration_count;
	const char* legacy = getenv("LSF_LEGACY_KERNELS");
	plan->allow_fast_kernels = !(legacy && legacy[0] == '1');
	const char* stage1 = getenv("LSF_STAGE1_VARIANT");
	if (stage1 && stage1[0] == '1') plan->stage1_variant = 1;
	plan->lane_xv = plan->tikhonov ? 2 : 4;
	const char* split = getenv("LSF_SPLIT_X");
	if (split && split[0] == '0') plan->split_x = false;
	const char* chunk_a = getenv("LSF_XCHUNK_A");
	if (chunk_a && atoi(chunk_a) > 0) plan->x_chunk_stage1 = atoi(chunk_a);
	const char* chunk_b = getenv("LSF_XCHUNK_B");
	if (chunk_b && atoi(chunk_b) > 0) plan->x_chunk_filter = atoi(chunk_b);
	const char* tma = getenv("LSF_TMA");
	if (tma && tma[0] == '0') plan->tma = false;
	const char* pair_ty = getenv("LSF_PAIR_TY");
	if (pair_ty) plan->pair_tile_y = atoi(pair_ty);
	const char* chunk_t = getenv("LSF_XCHUNK_T");
	if (chunk_t && atoi(chunk_t) > 0) plan->x_chunk_tma = atoi(chunk_t);
	const char* chunk_y = getenv("LSF_YCHUNK_T");
	if (chunk_y && atoi(chunk_y) > 0) plan->y_chunk_tma = atoi(chunk_y);
	const char* xv = getenv("LSF_LANE_XV");
	if (xv && (atoi(xv) == 1 || atoi(xv) == 2 || atoi(xv) == 4 || atoi(xv) == 8)) plan->lane_xv = atoi(xv);
	return LSF_OK;
}

inline bool aligned16(const void* p) {
	return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

struct LevelState {
	Grid3 g;
	const float4* pack = nullptr;
	const float* canonical = nullptr;
	float* warp = nullptr;       // planes
	float* g_post = nullptr;     // planes: gradient after the iteration (g_prev of the next one)
	float* scratch_a = nullptr;  // planes
	float* scratch_b = nullptr;  // planes
	unsigned* max_sq_bits = nullptr;
	// slab decomposition (see HierIterArgs): defaults = whole volume
	bool slab = false;
	int x_begin = 0, x_end = 0, x_origin = 0, X_global = 0;
	int pack_X = 0, pack_origin = 0, pack_interior_low = 0, pack_interior_high = 0;
	int* violation = nullptr;
	TmaMaps maps;                // tensor maps of this level's warp / canonical / gradient planes (encoded on first use)
	TmaMaps maps_alt;            // the same with the gradient's other ping-pong buffer (iterations without a Sobolev kernel)
	                             // or the warp's other ping-pong buffer (deferred update)
	float* scratch_h = nullptr;  // planes: slab mode, the gradient after the axis-0 pass (fast filter phase)
	// batch of pairs (HierIterArgs::batch_X): g.X = pairs * batch_X planes
	int batch_pairs = 0, batch_X = 0, batch_slot_stride = 0;
	long long batch_pack_stride = 0;
	float* warp_alt = nullptr;   // planes: second warp buffer. Non-null = deferred warp update (Tikhonov + Sobolev kernel):
	                             // iteration i reads warp (i even) / warp_alt (i odd) and writes the other one; after E
	                             // executed iterations finish_deferred() leaves the final warp in `warp`
};

// Enqueues one iteration; returns the number of kernel launches (negative: error status).
// `events` (optional, 5 entries): recorded before the first and after every kernel, for per-stage timing.
// `phase`: 0 = whole iteration; 1 = stage 1 only, 2 = filter stage only (slab mode: the halo exchange sits between).
int enqueue_iteration(const Plan3& plan, LevelState& s, int iteration, bool check_convergence, cudaStream_t stream,
		cudaEvent_t* events = nullptr, int phase = 0) {
	auto mark = [&](int i) {
		if (events) cudaEventRecord(events[i], stream);
	};
	mark(0);
	g_last_path = LSF_PATH_OTHER;
	HierIterArgs a;
	a.pack = s.pack;
	a.canonical = s.canonical;
	a.warp = s.warp;
	a.warp_out = s.warp;
	a.g_prev = s.g_post;
	a.g = s.g;
	a.amplifier = plan.amplifier;
	a.strength = plan.strength;
	a.rate = plan.rate;
	a.threshold = plan.threshold;
	a.max_sq_bits = s.max_sq_bits;
	a.iteration = iteration;
	a.check_convergence = check_convergence ? 1 : 0;
	whole_volume(a);
	if (s.batch_pairs > 0) {
		a.batch_X = s.batch_X;
		a.batch_pack_stride = s.batch_pack_stride;
		a.batch_slot_stride = s.batch_slot_stride;
		a.pack_X = s.batch_X;
	}
	const int volumes = s.batch_pairs > 0 ? s.batch_pairs : 1;       // tiles of the wave model: per plane x volumes
	const int planes_per_volume = s.batch_pairs > 0 ? s.batch_X : s.g.X;
	if (s.slab) {
		a.x_begin = s.x_begin;
		a.x_end = s.x_end;
		a.x_origin = s.x_origin;
		a.X_global = s.X_global;
		a.pack_X = s.pack_X;
		a.pack_origin = s.pack_origin;
		a.pack_interior_low = s.pack_interior_low;
		a.pack_interior_high = s.pack_interior_high;
		a.violation = s.violation;
	}
	const int x_begin = a.x_begin, x_end = a.x_end;
	dim3 grid = grid3(s.g), block = block3();
	// stage-1 variants: 0 = first generation (one voxel per thread, 64-bit indices), 1 = 4 z-voxels per thread
	// (128-bit loads), 2 = lane-contiguous x-marching (default: best L1 behaviour, see profiles/)
	const long long pack_count = (long long) (a.pack_X + 4) * (s.g.Y + 4) * (s.g.Z + 4);
	const bool small_indices = s.g.N * 3 < (1ll << 31) && pack_count < (1ll << 31);
	int variant = 0;
	if (plan.allow_fast_kernels && small_indices) {
		variant = 2;
		if (plan.stage1_variant == 1 && s.g.Z % 4 == 0 && aligned16(s.canonical) && aligned16(s.warp)
				&& aligned16(s.g_post) && aligned16(s.scratch_a)) variant = 1;
	}
	if (variant == 1) launch_shape_v4(s.g, &grid, &block);
	if (variant == 2) launch_shape_lane(s.g, x_end - x_begin, plan.lane_xv, &grid, &block);
#define LSF_LAUNCH_STAGE1(TIK, FUSE)                                                                      \
	do {                                                                                                  \
		if (variant == 2 && plan.lane_xv == 1) k_hier_gradient3d_lane<TIK, FUSE, 1> <<<counted(grid), block, 0, stream>>>(a); \
		else if (variant == 2 && plan.lane_xv == 2) k_hier_gradient3d_lane<TIK, FUSE, 2> <<<counted(grid), block, 0, stream>>>(a); \
		else if (variant == 2 && plan.lane_xv == 8) k_hier_gradient3d_lane<TIK, FUSE, 8> <<<counted(grid), block, 0, stream>>>(a); \
		else if (variant == 2) k_hier_gradient3d_lane<TIK, FUSE, 4> <<<counted(grid), block, 0, stream>>>(a); \
		else if (variant == 1) k_hier_gradient3d_v4<TIK, FUSE> <<<counted(grid), block, 0, stream>>>(a);       \
		else k_hier_gradient3d<TIK, FUSE> <<<counted(grid), block, 0, stream>>>(a);                          \
	} while (0)
	if (!plan.use_kernel) {
		if (phase == 2) return 0;
		if (plan.tma && variant == 2 && (!s.slab || plan.slab_fast) && tma_supported(s.g, s.warp, s.canonical, s.g_post)
				&& (!plan.tikhonov || aligned16(s.scratch_a))) {
			// TMA-fed stage 1 with the warp update and the max-norm fused (one launch per iteration)
			g_last_path = LSF_PATH_TMA_STAGE1 | LSF_PATH_FUSED_UPDATE;
			const int tiles = (int) (div_up(s.g.Z, 32) * div_up(s.g.Y, 8)) * volumes;
			const int planes = s.batch_pairs > 0 ? planes_per_volume : x_end - x_begin;
			const int chunk_x = plan.x_chunk_tma > 0 ? std::min(plan.x_chunk_tma, planes) : marching_chunk(planes, tiles, 0, 3);
			int status;
			if (plan.tikhonov) {
				// the gradient ping-pongs between g_post and scratch_a: one set of tensor maps per direction
				a.g_out = s.scratch_a;
				TmaMaps& maps = (s.maps.key[2] == nullptr || s.maps.key[2] == a.g_prev) ? s.maps : s.maps_alt;
				status = s.slab ? launch_stage1_fused_update<true, true>(maps, a, chunk_x, stream)
						: launch_stage1_fused_update<true>(maps, a, chunk_x, stream);
				std::swap(s.g_post, s.scratch_a);
			} else {
				a.g_out = nullptr;
				status = s.slab ? launch_stage1_fused_update<false, true>(s.maps, a, chunk_x, stream)
						: launch_stage1_fused_update<false>(s.maps, a, chunk_x, stream);
			}
			mark(1);
			return status < 0 ? status : 1;
		}
		if (plan.tikhonov) {
			a.g_out = s.scratch_a;
			LSF_LAUNCH_STAGE1(true, true);
			std::swap(s.g_post, s.scratch_a);
		} else {
			a.g_out = nullptr;
			LSF_LAUNCH_STAGE1(false, true);
		}
		mark(1);
		return 1;
	}
	if (plan.split_x && variant == 2 && !s.slab && phase == 0 && plan.taps.radius >= 1 && plan.taps.radius <= 3) {
		// second-generation cut: stage 1 + axis-0 pass, then axis-1/2 passes + update (kernels3d_split.cuh)
		float* filtered = plan.tikhonov ? s.g_post : nullptr;
		if (plan.tma && tma_supported(s.g, s.warp, s.canonical, s.g_post) && aligned16(s.scratch_a)) {
			const int tiles = (int) (div_up(s.g.Z, 32) * div_up(s.g.Y, 8)) * volumes;
			const int chunk_x = plan.x_chunk_tma > 0 ? std::min(plan.x_chunk_tma, planes_per_volume)
					: marching_chunk(planes_per_volume, tiles, 2 * plan.taps.radius, 3);
			const int filter_tiles = (int) (div_up(s.g.Z, 512) * s.g.X);
			const int chunk_y = plan.y_chunk_tma > 0 ? std::min(plan.y_chunk_tma, s.g.Y)
					: marching_chunk(s.g.Y, filter_tiles, 2 * plan.taps.radius, 6);
			if (s.warp_alt != nullptr) {
				g_last_path = LSF_PATH_TMA_STAGE1 | LSF_PATH_DEFERRED_UPDATE
						| (ymarch3_supported(s.g, s.scratch_a, s.g_post, s.g_post) ? LSF_PATH_YMARCH3 : LSF_PATH_YMARCH2);
				a.warp = iteration % 2 == 0 ? s.warp : s.warp_alt;
				a.warp_out = iteration % 2 == 0 ? s.warp_alt : s.warp;
				TmaMaps& maps = iteration % 2 == 0 ? s.maps : s.maps_alt;
				int status;
				switch (plan.taps.radius) {
				case 1:
					status = launch_iteration_deferred<1>(maps, a, plan.taps, s.scratch_a, filtered, chunk_x, chunk_y, stream, events);
					break;
				case 2:
					status = launch_iteration_deferred<2>(maps, a, plan.taps, s.scratch_a, filtered, chunk_x, chunk_y, stream, events);
					break;
				default:
					status = launch_iteration_deferred<3>(maps, a, plan.taps, s.scratch_a, filtered, chunk_x, chunk_y, stream, events);
					break;
				}
				return status < 0 ? status : 2;
			}
			g_last_path = LSF_PATH_TMA_STAGE1 | (ymarch3_supported(s.g, s.scratch_a, s.g_post, s.warp) ? LSF_PATH_YMARCH3 : LSF_PATH_YMARCH2);
			int status;
			switch (plan.taps.radius) {
			case 1:
				status = launch_iteration_v4<1>(plan.tikhonov, s.maps, a, plan.taps, s.scratch_a, filtered, s.warp, chunk_x, chunk_y, plan.pair_tile_y, stream, events);
				break;
			case 2:
				status = launch_iteration_v4<2>(plan.tikhonov, s.maps, a, plan.taps, s.scratch_a, filtered, s.warp, chunk_x, chunk_y, plan.pair_tile_y, stream, events);
				break;
			default:
				status = launch_iteration_v4<3>(plan.tikhonov, s.maps, a, plan.taps, s.scratch_a, filtered, s.warp, chunk_x, chunk_y, plan.pair_tile_y, stream, events);
				break;
			}
			return status < 0 ? status : 2;
		}
		const int chunk_a = std::min(plan.x_chunk_stage1, s.g.X), chunk_b = std::min(plan.x_chunk_filter, s.g.X);
		switch (plan.taps.radius) {
		case 1:
			launch_split_iteration<1>(plan.tikhonov, a, plan.taps, s.scratch_a, filtered, s.warp, chunk_a, chunk_b, stream, events);
			break;
		case 2:
			launch_split_iteration<2>(plan.tikhonov, a, plan.taps, s.scratch_a, filtered, s.warp, chunk_a, chunk_b, stream, events);
			break;
		default:
			launch_split_iteration<3>(plan.tikhonov, a, plan.taps, s.scratch_a, filtered, s.warp, chunk_a, chunk_b, stream, events);
			break;
		}
		return 2;
	}
	if (s.slab && plan.slab_fast && plan.tma && variant == 2 && (phase == 1 || s.scratch_h != nullptr) && plan.taps.radius >= 1
			&& plan.taps.radius <= 3 && tma_supported(s.g, s.warp, s.canonical, s.g_post) && aligned16(s.scratch_a)
			&& ymarch2_supported(s.g, s.scratch_h, s.g_post, s.warp)) {
		// slab mode, fourth-generation kernels: phase 1 = TMA-fed stage 1 writing the unfiltered gradient of the own
		// planes, [halo exchange by the caller,] phase 2 = axis-0 marching kernel + paired y-marching kernel
		int launched = 0;
		if (phase != 2) {
			const int tiles = (int) (div_up(s.g.Z, 32) * div_up(s.g.Y, 8));
			a.g_out = s.scratch_a;
			a.warp_out = nullptr;
			const int chunk_x = marching_chunk(x_end - x_begin, tiles, 0, 3);
			const int status = plan.tikhonov ? launch_stage1_fused_update<true, true>(s.maps, a, chunk_x, stream)
					: launch_stage1_fused_update<false, true>(s.maps, a, chunk_x, stream);
			if (status < 0) return status;
			launched++;
		}
		if (phase == 1) return launched;
		float* filtered = s.g_post;
		if (plan.taps.radius == 1) launch_slab_filter<1>(plan.taps, a, s.scratch_a, s.scratch_h, filtered, s.warp, stream);
		else if (plan.taps.radius == 2) launch_slab_filter<2>(plan.taps, a, s.scratch_a, s.scratch_h, filtered, s.warp, stream);
		else launch_slab_filter<3>(plan.taps, a, s.scratch_a, s.scratch_h, filtered, s.warp, stream);
		return launched + 2;
	}
	a.g_out = s.scratch_a;
	if (phase != 2) {
		if (plan.tikhonov) LSF_LAUNCH_STAGE1(true, false);
		else LSF_LAUNCH_STAGE1(false, false);
	}
#undef LSF_LAUNCH_STAGE1
	mark(1);
	if (phase == 1) return 1;
	// stage 2: the three filter passes + update + max-norm in ONE kernel for the usual 3/5/7-tap kernels
	if (plan.allow_fast_kernels && (plan.taps.radius >= 1 && plan.taps.radius <= 3)) {
		float* filtered = plan.tikhonov ? s.g_post : nullptr;  // only the Tikhonov term reads g of the last iteration
		const int check = a.check_convergence;
		launch_fused_filter_any(plan.taps, plan.rate, plan.threshold, s.g, s.scratch_a, filtered, s.warp, s.max_sq_bits,
				iteration, check, stream, x_begin, x_end);
		mark(2);
		return phase == 2 ? 1 : 2;
	}
	grid = grid3(s.g);
	block = block3();
	ConvArgs c;
	c.g = s.g;
	c.taps = plan.taps;
	c.rate = plan.rate;
	c.threshold = plan.threshold;
	c.max_sq_bits = s.max_sq_bits;
	c.iteration = iteration;
	c.check_convergence = a.check_convergence;
	c.channels = 3;
	c.warp = s.warp;
	c.in = s.scratch_a;
	c.out = s.scratch_b;
	k_convolve_axis3d<0, false> <<<counted(grid), block, 0, stream>>>(c);
	mark(2);
	c.in = s.scratch_b;
	c.out = s.scratch_a;
	k_convolve_axis3d<1, false> <<<counted(grid), block, 0, stream>>>(c);
	mark(3);
	c.in = s.scratch_a;
	c.out = s.g_post;
	k_convolve_axis3d<2, true> <<<counted(grid), block, 0, stream>>>(c);
	mark(4);
	return 4;
}

template<typename Access>
void launch_downsample(bool linear, Access access, const Grid3& src, const Grid3& dst, cudaStream_t stream) {
	if (linear) k_downsample_linear3d<Access> <<<counted(grid3(dst)), block3(), 0, stream>>>(access, src, dst);
	else k_downsample_average3d<Access> <<<counted(grid3(dst)), block3(), 0, stream>>>(access, src, dst);
}

int build_pack_pyramid(Arena& arena, const Plan3& plan, const float* live_dev, const float* canonical_dev,
		std::vector<float4*>& packs, std::vector<const float*>& canonicals, cudaStream_t stream, int first_level = 0) {
	const int L = plan.level_count;
	packs.assign(L, nullptr);
	canonicals.assign(L, nullptr);
	const float4 border = make_float4(1.0f, 0.0f, 0.0f, 0.0f);  // TSDF -> 1 (field_warping.tpp:29-36), gradient -> 0
	for (int level = L - 1; level >= first_level; level--) {
		const Grid3& g = plan.level_grid[level];
		LSF_TRY(arena.alloc(&packs[level], (size_t) g.padded_count()));
		k_fill4<<<counted(div_up(g.padded_count(), 256)), 256, 0, stream>>>(packs[level], g.padded_count(), border);
		if (level == L - 1) {
			k_gradient_pack3d<<<counted(grid3(g)), block3(), 0, stream>>>(live_dev, packs[level], g);
			canonicals[level] = canonical_dev;
		} else {
			const Grid3& src = plan.level_grid[level + 1];
			launch_downsample(plan.linear, PackAccess { packs[level + 1], packs[level] }, src, g, stream);
			float* canonical_level = nullptr;
			LSF_TRY(arena.alloc(&canonical_level, (size_t) g.N));
			launch_downsample(plan.linear, PlainAccess { canonicals[level + 1], canonical_level }, src, g, stream);
			canonicals[level] = canonical_level;
		}
	}
	LSF_CUDA(cudaGetLastError());
	return LSF_OK;
}

// live value of every voxel of a pack level (the .x lane of the padded float4 grid) as a plain scalar field
static __global__ void k_unpack_live3d(const float4* __restrict__ pack, float* __restrict__ out, Grid3 g) {
	LSF_VOXEL_3D(g);
	if (!in_grid) return;
	out[idx] = pack[g.padded_index(x, y, z)].x;
}

void fill_report_statistics(lsf_level_report& r, const lsf_warp_delta_statistics_t& w,
		const lsf_tsdf_difference_statistics_t& d) {
	r.warp_ratio_above_min_threshold = w.ratio_above_min_threshold;
	r.warp_length_min = w.length_min;
	r.warp_length_max = w.length_max;
	r.warp_length_mean = w.length_mean;
	r.warp_length_std = w.length_standard_deviation;
	for (int i = 0; i < 3; i++) r.warp_longest_location[i] = w.longest_warp_location[i];
	r.warp_is_largest_below_min_threshold = w.is_largest_below_min_threshold;
	r.warp_is_largest_above_max_threshold = w.is_largest_above_max_threshold;
	r.diff_min = d.difference_min;
	r.diff_max = d.difference_max;
	r.diff_mean = d.difference_mean;
	r.diff_std = d.difference_standard_deviation;
	for (int i = 0; i < 3; i++) r.diff_biggest_location[i] = d.biggest_difference_location[i];
}

// diff = warped live - canonical and the data-term gradient (resampled live gradient * diff) of the current warp
// (reference optimizer.tpp:186-194), for the telemetry path
static __global__ void k_telemetry_terms3d(const float4* __restrict__ pack, const float* __restrict__ canonical,
		const float* __restrict__ warp, float* __restrict__ diff, float* __restrict__ data_planes, Grid3 g) {
	LSF_VOXEL_3D(g);
	if (!in_grid) return;
	const float4 s = gather4(pack, g, x, y, z, warp[idx], warp[g.N + idx], warp[2 * g.N + idx]);
	const float d = s.x - canonical[idx];
	diff[idx] = d;
	data_planes[idx] = s.y * d;
	data_planes[g.N + idx] = s.z * d;
	data_planes[2 * g.N + idx] = s.w * d;
}

// One level with per-iteration telemetry (see lsf_iteration_sink): one iteration at a time, first-in-line kernels of the
// production path for the iteration itself (no deferred update: the warp after every iteration is wanted).
int run_level_with_telemetry(const Plan3& plan, LevelState& s, int level, const lsf_iteration_sink* sink,
		TelemetryScratch& t, cudaStream_t stream, int* executed_out, float* last_max_out) {
	const long long N = s.g.N;
	const int dims[3] = { s.g.X, s.g.Y, s.g.Z };
	lsf_iteration_record record;
	std::memset(&record, 0, sizeof(record));
	record.level = level;
	for (int i = 0; i < 3; i++) record.dims[i] = dims[i];
	k_unpack_live3d<<<counted(grid3(s.g)), block3(), 0, stream>>>(s.pack, t.live_level, s.g);
	if (level == 0 && sink->want_fields) {
		// reference optimizer_with_telemetry.tpp:90-99: the first frame of level 0 holds the live level and zero fields
		LSF_CUDA(cudaMemsetAsync(t.data_planes, 0, (size_t) N * 3 * sizeof(float), stream));
		if (plan.tikhonov) LSF_CUDA(cudaMemsetAsync(t.tikhonov_planes, 0, (size_t) N * 3 * sizeof(float), stream));
		record.iteration = -1;
		LSF_TRY(telemetry_fields(t, N, 3, t.data_planes, plan.tikhonov, &record, stream));
		sink->callback(sink->user, &record);
	}
	int executed = 0;
	float last_max = FLT_MAX;
	unsigned bits = 0;
	for (int it = 0; it < plan.max_iterations; it++) {
		if (last_max < plan.threshold) break;  // reference optimizer.tpp:149,166-171
		k_telemetry_terms3d<<<counted(grid3(s.g)), block3(), 0, stream>>>(s.pack, s.canonical, s.warp, t.diff, t.data_planes, s.g);
		if (plan.tikhonov)
			k_laplacian_planes3d<<<counted(grid3(s.g)), block3(), 0, stream>>>(s.g_post, t.tikhonov_planes, 3, s.g);
		record.iteration = it;
		if (sink->want_statistics) LSF_TRY(telemetry_statistics(t, N, 3, dims, s.g_post, plan.tikhonov, &record, stream));
		LSF_TRY(enqueue_iteration(plan, s, it, false, stream));
		LSF_CUDA(cudaMemcpyAsync(&bits, s.max_sq_bits + it, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
		LSF_CUDA(cudaStreamSynchronize(stream));
		float sq;
		std::memcpy(&sq, &bits, sizeof(float));
		last_max = std::sqrt(sq);
		executed = it + 1;
		record.max_update_length = last_max;
		if (sink->want_fields) LSF_TRY(telemetry_fields(t, N, 3, s.warp, plan.tikhonov, &record, stream));
		sink->callback(sink->user, &record);
	}
	*executed_out = executed;
	*last_max_out = last_max;
	return LSF_OK;
}

// Deferred warp update (k_hier_stage1_tma APPLY): usable when the level runs the TMA generation with both the Tikhonov
// term and a Sobolev kernel, on the whole volume.
bool deferred_update_applies(const Plan3& plan, const LevelState& s) {
	const char* e = getenv("LSF_DEFER");  // A/B: LSF_DEFER=0 keeps the update in the filter kernel
	if (e && e[0] == '0') return false;
	const long long pack_count = (long long) (s.g.X + 4) * (s.g.Y + 4) * (s.g.Z + 4);
	return plan.tikhonov && plan.use_kernel && plan.allow_fast_kernels && plan.split_x && plan.tma && !s.slab
			&& plan.taps.radius >= 1 && plan.taps.radius <= 3 && plan.pair_tile_y == 0 && s.g.N * 3 < (1ll << 31)
			&& pack_count < (1ll << 31) && tma_supported(s.g, s.warp, s.canonical, s.g_post) && aligned16(s.scratch_a)
			&& deferred_update_supported(s.g, s.scratch_a, s.g_post);
}

// After `executed` iterations of a deferred level: apply the last iteration's pending update, result in s.warp.
void finish_deferred(const Plan3& plan, LevelState& s, int executed, cudaStream_t stream) {
	if (s.warp_alt == nullptr) return;
	const float* current = executed % 2 == 0 ? s.warp : s.warp_alt;
	const long long count = s.g.N * 3;
	k_apply_update3d<<<counted(div_up(count, 256)), 256, 0, stream>>>(current, s.g_post, s.warp, plan.rate, count);
}

int optimize_device(const Plan3& plan, const float* canonical_dev, const float* live_dev, float* warp_out_dev,
		lsf_level_report* reports, int collect_reports, lsf_iteration_capture* capture, float* capture_dev,
		cudaStream_t stream, const lsf_iteration_sink* sink = nullptr) {
	Arena arena(stream);
	const int L = plan.level_count;
	const Grid3& finest = plan.level_grid[L - 1];
	std::vector<float4*> packs;
	std::vector<const float*> canonicals;
	LSF_TRY(build_pack_pyramid(arena, plan, live_dev, canonical_dev, packs, canonicals, stream));

	float *warp_a, *warp_b, *g_post, *scratch_a = nullptr, *scratch_b = nullptr;
	unsigned* max_sq_bits;
	LSF_TRY(arena.alloc(&warp_a, (size_t) finest.N * 3));
	LSF_TRY(arena.alloc(&warp_b, (size_t) finest.N * 3 / 8 + 16));
	LSF_TRY(arena.alloc(&g_post, (size_t) finest.N * 3));
	if (plan.use_kernel || plan.tikhonov) LSF_TRY(arena.alloc(&scratch_a, (size_t) finest.N * 3));
	if (plan.use_kernel) LSF_TRY(arena.alloc(&scratch_b, (size_t) finest.N * 3));
	const int slot_count = std::max(plan.max_iterations, 1);
	LSF_TRY(arena.alloc(&max_sq_bits, (size_t) slot_count));
	PollState& poll = poll_state();
	LSF_TRY(poll.reserve((size_t) slot_count));
	unsigned* host_bits = poll.host_bits;
	const char* pipeline_env = getenv("LSF_PIPELINE_POLL");  // A/B: 0 = drain the stream at every poll
	const bool pipelined = !(pipeline_env && pipeline_env[0] == '0');
	float* warp_pong = nullptr;  // second warp buffer of the deferred update
	if (plan.tikhonov && plan.use_kernel && sink == nullptr) LSF_TRY(arena.alloc(&warp_pong, (size_t) finest.N * 3));
	TelemetryScratch telemetry;
	if (sink != nullptr) LSF_TRY(telemetry.allocate(arena, (size_t) finest.N, 3, plan.tikhonov, sink->want_fields != 0));

	// the finest level lives in warp_a (3N floats), the level below it in warp_b (3N/8), and so on alternating
	float* warp_current = ((L - 1) % 2 == 0) ? warp_a : warp_b;
	float* warp_next = ((L - 1) % 2 == 0) ? warp_b : warp_a;
	LSF_CUDA(cudaMemsetAsync(warp_current, 0, (size_t) plan.level_grid[0].N * 3 * sizeof(float), stream));
	if (capture) capture->count = 0;

	for (int level = 0; level < L; level++) {
		LevelState s;
		s.g = plan.level_grid[level];
		s.pack = packs[level];
		s.canonical = canonicals[level];
		s.warp = warp_current;
		s.g_post = g_post;
		s.scratch_a = scratch_a;
		s.scratch_b = scratch_b;
		s.max_sq_bits = max_sq_bits;
		// reference optimizer.tpp:142-143: gradient = 0 at the start of every level
		LSF_CUDA(cudaMemsetAsync(g_post, 0, (size_t) s.g.N * 3 * sizeof(float), stream));
		LSF_CUDA(cudaMemsetAsync(max_sq_bits, 0, (size_t) slot_count * sizeof(unsigned), stream));
		const bool capturing = capture && capture->level == level && capture_dev != nullptr;
		if (warp_pong != nullptr && deferred_update_applies(plan, s)) s.warp_alt = warp_pong;
		// Per-iteration captures (reference optimizer_with_telemetry.tpp:153-159) on the deferred path: the warp after
		// iteration i is materialised by stage 1 of iteration i + 1 (its warp_out planes), so capture slot i is copied
		// after iteration i + 1 has been enqueued; the slot of the last executed iteration is filled after
		// finish_deferred(). Kernels of iterations beyond the converged one return early: their slots are never counted.
		auto capture_slot = [&](const float* planes, int slot) {
			k_planes_to_aos<<<counted(div_up(s.g.N, 256)), 256, 0, stream>>>(planes, capture_dev + (size_t) slot * s.g.N * 3,
					s.g.N, 3);
		};

		// The termination test (reference optimizer.tpp:166-171) runs on the device at the head of every kernel; the host
		// only has to learn the iteration count. Chunks of POLL_CHUNK iterations are enqueued one ahead of the chunk
		// whose slots the host is waiting for (pinned staging buffer, one event per chunk), so the stream never drains
		// while the host looks at the results. Kernels enqueued beyond the converged iteration return at once.
		int executed = 0;      // iterations known to have run
		int enqueued = 0;
		bool converged = false;
		float last_max = FLT_MAX;
		auto enqueue_chunk = [&](int begin, int end, cudaEvent_t done) -> int {
			for (int it = begin; it < end; it++) {
				LSF_TRY(enqueue_iteration(plan, s, it, true, stream));
				if (capturing && s.warp_alt == nullptr && it < capture->max_iterations) capture_slot(s.warp, it);
				if (capturing && s.warp_alt != nullptr && it >= 1 && it - 1 < capture->max_iterations)
					capture_slot(it % 2 == 0 ? s.warp_alt : s.warp, it - 1);  // the buffer stage 1 of iteration `it` wrote
			}
			LSF_CUDA(cudaGetLastError());
			LSF_CUDA(cudaMemcpyAsync(host_bits + begin, max_sq_bits + begin, (size_t) (end - begin) * sizeof(unsigned),
					cudaMemcpyDeviceToHost, stream));
			LSF_CUDA(cudaEventRecord(done, stream));
			return LSF_OK;
		};
		int pending_begin = 0, pending_end = std::min(plan.max_iterations, POLL_CHUNK);
		int parity = 0;
		if (sink != nullptr) {
			LSF_TRY(run_level_with_telemetry(plan, s, level, sink, telemetry, stream, &executed, &last_max));
			pending_end = 0;
		}
		if (pending_end > 0) LSF_TRY(enqueue_chunk(0, pending_end, poll.events[parity]));
		enqueued = pending_end;
		while (!converged && pending_begin < pending_end) {
			// keep one more chunk in flight behind the one being polled
			const int next_end = std::min(plan.max_iterations, enqueued + POLL_CHUNK);
			const bool has_next = pipelined && next_end > enqueued;
			if (has_next) LSF_TRY(enqueue_chunk(enqueued, next_end, poll.events[parity ^ 1]));
			LSF_CUDA(cudaEventSynchronize(poll.events[parity]));
			for (int it = pending_begin; it < pending_end; it++) {
				float sq;
				std::memcpy(&sq, &host_bits[it], sizeof(float));
				last_max = std::sqrt(sq);
				executed = it + 1;
				if (last_max < plan.threshold) {  // reference optimizer.tpp:166-171
					converged = true;
					break;
				}
			}
			pending_begin = pending_end;
			if (has_next) {
				pending_end = next_end;
				enqueued = next_end;
				parity ^= 1;
			} else if (!converged && !pipelined && enqueued < plan.max_iterations) {
				pending_end = next_end;
				LSF_TRY(enqueue_chunk(enqueued, next_end, poll.events[parity]));
				enqueued = next_end;
			}
		}
		finish_deferred(plan, s, executed, stream);
		if (capturing && s.warp_alt != nullptr && executed >= 1 && executed - 1 < capture->max_iterations)
			capture_slot(s.warp, executed - 1);
		if (reports) {
			lsf_level_report& r = reports[level];
			std::memset(&r, 0, sizeof(r));
			r.iteration_count = executed;
			r.iteration_limit_reached = executed >= plan.max_iterations;
			r.max_update_length = last_max;
			r.dims[0] = s.g.X;
			r.dims[1] = s.g.Y;
			r.dims[2] = s.g.Z;
			if (collect_reports) {
				// reference optimizer_with_telemetry.tpp:107-124: statistics of the level's warp field over the band
				// union of (canonical level, live level) and |live level - canonical level|
				float* live_level;
				LSF_TRY(arena.alloc(&live_level, (size_t) s.g.N));
				k_unpack_live3d<<<counted(grid3(s.g)), block3(), 0, stream>>>(s.pack, live_level, s.g);
				lsf_warp_delta_statistics_t w;
				lsf_tsdf_difference_statistics_t d;
				LSF_TRY(statistics_on_device(3, r.dims, s.warp, s.g.N, 1, s.canonical, live_level, plan.threshold, FLT_MAX, &w,
						&d, arena, stream));
				fill_report_statistics(r, w, d);
			}
		}
		if (capturing) capture->count = std::min(executed, capture->max_iterations);
		if (level != L - 1) {
			// reference optimizer.tpp:124-126: prolong the warp (values are NOT doubled)
			const Grid3& dg = plan.level_grid[level + 1];
			if (plan.linear) k_upsample_linear3d<<<counted(grid3(dg)), block3(), 0, stream>>>(warp_current, warp_next, 3, s.g, dg);
			else k_upsample_nearest3d<<<counted(grid3(dg)), block3(), 0, stream>>>(warp_current, warp_next, 3, s.g, dg);
			std::swap(warp_current, warp_next);
		}
	}
	k_planes_to_aos<<<counted(div_up(finest.N, 256)), 256, 0, stream>>>(warp_current, warp_out_dev, finest.N, 3);
	LSF_CUDA(cudaGetLastError());
	return LSF_OK;
}

// Pending update of a deferred level for a batch of pairs: pair p stopped after executed[p] iterations, so its warp
// before the last update sits in `even` (executed[p] even) or `odd`; out = that - g * rate (out may alias even).
static __global__ void k_apply_update3d_batch(const float* __restrict__ even, const float* __restrict__ odd,
		const float* __restrict__ g, float* __restrict__ out, float rate, long long pair_voxels, long long batch_voxels,
		const int* __restrict__ executed) {
	const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= 3 * batch_voxels) return;
	const int pair = (int) ((i % batch_voxels) / pair_voxels);
	const float w = (executed[pair] % 2 == 0) ? even[i] : odd[i];
	out[i] = w - g[i] * rate;
}

// Can the batched level loop run this plan on `pairs` pairs at once? (else the caller loops over the pairs)
bool batch_supported(const Plan3& plan, int pairs) {
	const char* e = getenv("LSF_BATCH");  // A/B: LSF_BATCH=0 optimises the pairs one after the other
	if (e && e[0] == '0') return false;
	if (pairs < 2 || plan.linear || !plan.allow_fast_kernels || !plan.tma || !plan.split_x || plan.pair_tile_y != 0) return false;
	if (plan.use_kernel && (plan.taps.radius < 1 || plan.taps.radius > 3)) return false;
	for (int level = 0; level < plan.level_count; level++) {
		const Grid3& g = plan.level_grid[level];
		if (g.Z % 4 != 0 || g.N % 2 != 0) return false;
		if ((long long) pairs * g.N * 3 >= (1ll << 31)) return false;
		if ((long long) (g.X + 4) * (g.Y + 4) * (g.Z + 4) >= (1ll << 31)) return false;
	}
	return true;
}

// Batched optimize() (reference loop run_hierarchical_optimizer3d_multipair.py:403-406 over independent pairs; SURVEY
// 8e): the pairs advance in lockstep through the pyramid levels and every iteration kernel covers all of them
// (HierIterArgs::batch_X), so the launch-bound coarse levels cost one launch sequence for the whole batch. Termination
// is per pair: each pair has its own row of convergence slots, its kernels' blocks return once it has converged, and
// the level ends when every pair has. Results per pair are bit-identical to lsf_hier_optimize_3d.
int optimize_batch_device(const Plan3& plan, int pairs, const float* canonical_dev, const float* live_dev,
		float* warp_out_dev, int* iteration_counts, cudaStream_t stream) {
	Arena arena(stream);
	const int L = plan.level_count;
	const Grid3& finest = plan.level_grid[L - 1];
	const size_t P = (size_t) pairs;
	const float4 border = make_float4(1.0f, 0.0f, 0.0f, 0.0f);
	// pyramids: one padded pack per pair and level, canonical levels one pair after the other
	std::vector<float4*> packs(L, nullptr);
	std::vector<const float*> canonicals(L, nullptr);
	for (int level = L - 1; level >= 0; level--) {
		const Grid3& g = plan.level_grid[level];
		const long long padded = g.padded_count();
		LSF_TRY(arena.alloc(&packs[level], (size_t) padded * P));
		k_fill4<<<counted(div_up(padded * pairs, 256)), 256, 0, stream>>>(packs[level], padded * pairs, border);
		float* canonical_level = nullptr;
		if (level == L - 1) canonicals[level] = canonical_dev;
		else {
			LSF_TRY(arena.alloc(&canonical_level, (size_t) g.N * P));
			canonicals[level] = canonical_level;
		}
		for (int pair = 0; pair < pairs; pair++) {
			if (level == L - 1) {
				k_gradient_pack3d<<<counted(grid3(g)), block3(), 0, stream>>>(live_dev + (size_t) pair * g.N,
						packs[level] + (size_t) pair * padded, g);
			} else {
				const Grid3& src = plan.level_grid[level + 1];
				launch_downsample(false, PackAccess { packs[level + 1] + (size_t) pair * src.padded_count(),
						packs[level] + (size_t) pair * padded }, src, g, stream);
				launch_downsample(false, PlainAccess { canonicals[level + 1] + (size_t) pair * src.N,
						canonical_level + (size_t) pair * g.N }, src, g, stream);
			}
		}
	}
	LSF_CUDA(cudaGetLastError());

	float *warp_a, *warp_b, *g_post, *scratch_a = nullptr, *warp_pong = nullptr;
	unsigned* max_sq_bits;
	int* executed_dev;
	LSF_TRY(arena.alloc(&warp_a, (size_t) finest.N * 3 * P));
	LSF_TRY(arena.alloc(&warp_b, ((size_t) finest.N * 3 / 8 + 16) * P));
	LSF_TRY(arena.alloc(&g_post, (size_t) finest.N * 3 * P));
	if (plan.use_kernel || plan.tikhonov) LSF_TRY(arena.alloc(&scratch_a, (size_t) finest.N * 3 * P));
	if (plan.tikhonov && plan.use_kernel) LSF_TRY(arena.alloc(&warp_pong, (size_t) finest.N * 3 * P));
	const int slot_count = std::max(plan.max_iterations, 1);
	LSF_TRY(arena.alloc(&max_sq_bits, (size_t) slot_count * P));
	LSF_TRY(arena.alloc(&executed_dev, P));
	PollState& poll = poll_state();
	LSF_TRY(poll.reserve((size_t) slot_count * P));
	unsigned* host_bits = poll.host_bits;  // [pair][slot]

	float* warp_current = ((L - 1) % 2 == 0) ? warp_a : warp_b;
	float* warp_next = ((L - 1) % 2 == 0) ? warp_b : warp_a;
	LSF_CUDA(cudaMemsetAsync(warp_current, 0, (size_t) plan.level_grid[0].N * 3 * P * sizeof(float), stream));
	std::vector<int> executed(P), done(P);
	std::vector<float> last_max(P);

	for (int level = 0; level < L; level++) {
		const Grid3& lg = plan.level_grid[level];
		LevelState s;
		s.g = Grid3(lg.X * pairs, lg.Y, lg.Z);
		s.batch_pairs = pairs;
		s.batch_X = lg.X;
		s.batch_pack_stride = lg.padded_count();
		s.batch_slot_stride = slot_count;
		s.pack = packs[level];
		s.canonical = canonicals[level];
		s.warp = warp_current;
		s.g_post = g_post;
		s.scratch_a = scratch_a;
		s.scratch_b = nullptr;
		s.max_sq_bits = max_sq_bits;
		LSF_CUDA(cudaMemsetAsync(g_post, 0, (size_t) s.g.N * 3 * sizeof(float), stream));
		LSF_CUDA(cudaMemsetAsync(max_sq_bits, 0, (size_t) slot_count * P * sizeof(unsigned), stream));
		LSF_REQUIRE(tma_supported(s.g, s.warp, s.canonical, s.g_post) && ymarch2_supported(s.g, s.scratch_a, s.g_post, s.warp),
				"internal: the batched level loop does not apply to level %d", level);
		if (warp_pong != nullptr && deferred_update_applies(plan, s)) s.warp_alt = warp_pong;
		std::fill(executed.begin(), executed.end(), 0);
		std::fill(done.begin(), done.end(), 0);
		std::fill(last_max.begin(), last_max.end(), FLT_MAX);
		int remaining = pairs;
		auto enqueue_chunk = [&](int begin, int end, cudaEvent_t finished) -> int {
			for (int it = begin; it < end; it++) LSF_TRY(enqueue_iteration(plan, s, it, true, stream));
			LSF_CUDA(cudaGetLastError());
			LSF_CUDA(cudaMemcpy2DAsync(host_bits + begin, (size_t) slot_count * sizeof(unsigned), max_sq_bits + begin,
					(size_t) slot_count * sizeof(unsigned), (size_t) (end - begin) * sizeof(unsigned), P,
					cudaMemcpyDeviceToHost, stream));
			LSF_CUDA(cudaEventRecord(finished, stream));
			return LSF_OK;
		};
		int pending_begin = 0, pending_end = std::min(plan.max_iterations, POLL_CHUNK), enqueued = pending_end, parity = 0;
		if (pending_end > 0) LSF_TRY(enqueue_chunk(0, pending_end, poll.events[parity]));
		while (remaining > 0 && pending_begin < pending_end) {
			const int next_end = std::min(plan.max_iterations, enqueued + POLL_CHUNK);
			const bool has_next = next_end > enqueued;
			if (has_next) LSF_TRY(enqueue_chunk(enqueued, next_end, poll.events[parity ^ 1]));
			LSF_CUDA(cudaEventSynchronize(poll.events[parity]));
			for (int pair = 0; pair < pairs; pair++) {
				if (done[pair]) continue;
				for (int it = pending_begin; it < pending_end; it++) {
					float sq;
					std::memcpy(&sq, &host_bits[(size_t) pair * slot_count + it], sizeof(float));
					last_max[pair] = std::sqrt(sq);
					executed[pair] = it + 1;
					if (last_max[pair] < plan.threshold) {  // reference optimizer.tpp:166-171
						done[pair] = 1;
						remaining--;
						break;
					}
				}
			}
			pending_begin = pending_end;
			if (has_next) {
				pending_end = next_end;
				enqueued = next_end;
				parity ^= 1;
			}
		}
		if (s.warp_alt != nullptr) {
			LSF_CUDA(cudaMemcpyAsync(executed_dev, executed.data(), P * sizeof(int), cudaMemcpyHostToDevice, stream));
			const long long count = s.g.N * 3;
			k_apply_update3d_batch<<<counted(div_up(count, 256)), 256, 0, stream>>>(s.warp, s.warp_alt, s.g_post, s.warp,
					plan.rate, lg.N, s.g.N, executed_dev);
			LSF_CUDA(cudaStreamSynchronize(stream));  // `executed` is re-used by the next level
		}
		if (iteration_counts) {
			for (int pair = 0; pair < pairs; pair++) iteration_counts[pair * LSF_MAX_LEVELS + level] = executed[pair];
		}
		if (level != L - 1) {
			const Grid3& dl = plan.level_grid[level + 1];
			const Grid3 dg(dl.X * pairs, dl.Y, dl.Z);
			// reference optimizer.tpp:124-126 (nearest: every voxel's parent lies in the same pair)
			k_upsample_nearest3d<<<counted(grid3(dg)), block3(), 0, stream>>>(warp_current, warp_next, 3, s.g, dg);
			std::swap(warp_current, warp_next);
		}
	}
	k_planes_to_aos<<<counted(div_up(finest.N * pairs, 256)), 256, 0, stream>>>(warp_current, warp_out_dev, finest.N * pairs, 3);
	LSF_CUDA(cudaGetLastError());
	return LSF_OK;
}

}  // namespace

}  // namespace lsf

using namespace lsf;

extern "C" int lsf_hier_optimize_3d(const lsf_hier_params* params, const float* canonical, const float* live, int X,
		int Y, int Z, float* warp_out, int memory_kind, lsf_level_report* reports, int collect_reports,
		lsf_iteration_capture* capture, void* stream_handle) {
	cudaStream_t stream = static_cast<cudaStream_t>(stream_handle);
	Plan3 plan;
	LSF_TRY(make_plan(params, X, Y, Z, &plan));
	LSF_REQUIRE(canonical && live && warp_out, "canonical, live and warp_out must not be NULL");
	const size_t N = (size_t) X * Y * Z;
	Arena arena(stream);
	const float *canonical_dev, *live_dev;
	LSF_TRY(to_device(arena, canonical, N, memory_kind, stream, &canonical_dev));
	LSF_TRY(to_device(arena, live, N, memory_kind, stream, &live_dev));
	float* out_dev = warp_out;
	if (memory_kind == LSF_HOST) LSF_TRY(arena.alloc(&out_dev, N * 3));
	float* capture_dev = nullptr;
	size_t capture_count = 0;
	if (capture && capture->level >= 0 && capture->level < plan.level_count && capture->max_iterations > 0
			&& capture->buffer) {
		capture_count = (size_t) capture->max_iterations * plan.level_grid[capture->level].N * 3;
		if (memory_kind == LSF_HOST) LSF_TRY(arena.alloc(&capture_dev, capture_count));
		else capture_dev = capture->buffer;
	}
	LSF_TRY(optimize_device(plan, canonical_dev, live_dev, out_dev, reports, collect_reports, capture, capture_dev, stream));
	if (memory_kind == LSF_HOST) {
		if (capture_dev) {
			const size_t used = (size_t) capture->count * plan.level_grid[capture->level].N * 3;
			LSF_CUDA(cudaMemcpyAsync(capture->buffer, capture_dev, used * sizeof(float), cudaMemcpyDeviceToHost, stream));
		}
		LSF_TRY(from_device(out_dev, warp_out, N * 3, LSF_HOST, stream));
	}
	return plan.level_count;
}

extern "C" int lsf_hier_optimize_3d_telemetry(const lsf_hier_params* params, const float* canonical, const float* live,
		int X, int Y, int Z, float* warp_out, int memory_kind, lsf_level_report* reports, int collect_reports,
		const lsf_iteration_sink* sink, void* stream_handle) {
	cudaStream_t stream = static_cast<cudaStream_t>(stream_handle);
	Plan3 plan;
	LSF_TRY(make_plan(params, X, Y, Z, &plan));
	LSF_REQUIRE(canonical && live && warp_out, "canonical, live and warp_out must not be NULL");
	LSF_REQUIRE(sink == nullptr || sink->callback != nullptr, "the iteration sink has no callback");
	const size_t N = (size_t) X * Y * Z;
	Arena arena(stream);
	const float *canonical_dev, *live_dev;
	LSF_TRY(to_device(arena, canonical, N, memory_kind, stream, &canonical_dev));
	LSF_TRY(to_device(arena, live, N, memory_kind, stream, &live_dev));
	float* out_dev = warp_out;
	if (memory_kind == LSF_HOST) LSF_TRY(arena.alloc(&out_dev, N * 3));
	LSF_TRY(optimize_device(plan, canonical_dev, live_dev, out_dev, reports, collect_reports, nullptr, nullptr, stream, sink));
	if (memory_kind == LSF_HOST) LSF_TRY(from_device(out_dev, warp_out, N * 3, LSF_HOST, stream));
	return plan.level_count;
}

extern "C" int lsf_debug_last_path(void) {
	return lsf::g_last_path.load();
}

extern "C" int lsf_hier_optimize_3d_batch(const lsf_hier_params* params, const float* canonical, const float* live,
		int pair_count, int X, int Y, int Z, float* warp_out, int memory_kind, int* iteration_counts,
		void* stream_handle) {
	cudaStream_t stream = static_cast<cudaStream_t>(stream_handle);
	LSF_REQUIRE(pair_count >= 0, "pair_count must be non-negative");
	Plan3 plan;
	LSF_TRY(make_plan(params, X, Y, Z, &plan));
	const size_t N = (size_t) X * Y * Z;
	if (pair_count == 0) return plan.level_count;
	LSF_REQUIRE(canonical && live && warp_out, "canonical, live and warp_out must not be NULL");
	if (iteration_counts) std::memset(iteration_counts, 0, sizeof(int) * (size_t) pair_count * LSF_MAX_LEVELS);
	// sub-batches: 32-bit voxel indices inside the kernels (3 * pairs * N < 2^31) and a bound on the scratch memory
	// (about 17 fields of N floats per pair; LSF_BATCH_BYTES overrides the 24 GB default)
	const char* budget_env = getenv("LSF_BATCH_BYTES");
	const double budget = budget_env ? atof(budget_env) : 24e9;
	long long group = std::min<long long>(((1ll << 31) - 1) / (3 * (long long) N), (long long) (budget / (17.0 * 4.0 * N)));
	group = std::max<long long>(1, std::min<long long>(group, pair_count));
	for (int first = 0; first < pair_count; first += (int) group) {
		const int pairs = (int) std::min<long long>(group, pair_count - first);
		const float* canonical_group = canonical + (size_t) first * N;
		const float* live_group = live + (size_t) first * N;
		float* warp_group = warp_out + (size_t) first * N * 3;
		int* counts_group = iteration_counts ? iteration_counts + (size_t) first * LSF_MAX_LEVELS : nullptr;
		if (batch_supported(plan, pairs)) {
			Arena arena(stream);
			const float *canonical_dev, *live_dev;
			LSF_TRY(to_device(arena, canonical_group, N * pairs, memory_kind, stream, &canonical_dev));
			LSF_TRY(to_device(arena, live_group, N * pairs, memory_kind, stream, &live_dev));
			float* out_dev = warp_group;
			if (memory_kind == LSF_HOST) LSF_TRY(arena.alloc(&out_dev, N * 3 * pairs));
			LSF_TRY(optimize_batch_device(plan, pairs, canonical_dev, live_dev, out_dev, counts_group, stream));
			if (memory_kind == LSF_HOST) LSF_TRY(from_device(out_dev, warp_group, N * 3 * pairs, LSF_HOST, stream));
			else LSF_CUDA(cudaStreamSynchronize(stream));  // the arena's blocks are released in stream order anyway
			continue;
		}
		// shapes or settings outside the batched kernels: one pair after the other
		std::vector<lsf_level_report> reports(LSF_MAX_LEVELS);
		for (int pair = 0; pair < pairs; pair++) {
			const int levels = lsf_hier_optimize_3d(params, canonical_group + pair * N, live_group + pair * N, X, Y, Z,
					warp_group + pair * N * 3, memory_kind, reports.data(), 0, nullptr, stream_handle);
			if (levels < 0) return levels;
			if (counts_group) {
				for (int level = 0; level < levels; level++)
					counts_group[pair * LSF_MAX_LEVELS + level] = reports[level].iteration_count;
			}
		}
	}
	return plan.level_count;
}

extern "C" int lsf_hier_iterate_3d(const lsf_hier_params* params, const float* canonical_dev, const float* live_dev,
		int X, int Y, int Z, int iterations, float* elapsed_ms, int* kernel_launches, float* stage_ms,
		void* stream_handle) {
	cudaStream_t stream = static_cast<cudaStream_t>(stream_handle);
	Plan3 plan;
	LSF_TRY(make_plan(params, X, Y, Z, &plan));
	LSF_REQUIRE(iterations > 0, "iterations must be positive");
	Arena arena(stream);
	const int L = plan.level_count;
	const Grid3& g = plan.level_grid[L - 1];
	std::vector<float4*> packs;
	std::vector<const float*> canonicals;
	LSF_TRY(build_pack_pyramid(arena, plan, live_dev, canonical_dev, packs, canonicals, stream, L - 1));
	LevelState s;
	s.g = g;
	s.pack = packs[L - 1];
	s.canonical = canonical_dev;
	LSF_TRY(arena.alloc(&s.warp, (size_t) g.N * 3));
	LSF_TRY(arena.alloc(&s.g_post, (size_t) g.N * 3));
	LSF_TRY(arena.alloc(&s.scratch_a, (size_t) g.N * 3));
	LSF_TRY(arena.alloc(&s.scratch_b, (size_t) g.N * 3));
	LSF_TRY(arena.alloc(&s.max_sq_bits, (size_t) iterations));
	LSF_CUDA(cudaMemsetAsync(s.warp, 0, (size_t) g.N * 3 * sizeof(float), stream));
	LSF_CUDA(cudaMemsetAsync(s.g_post, 0, (size_t) g.N * 3 * sizeof(float), stream));
	LSF_CUDA(cudaMemsetAsync(s.max_sq_bits, 0, (size_t) iterations * sizeof(unsigned), stream));
	if (deferred_update_applies(plan, s)) LSF_TRY(arena.alloc(&s.warp_alt, (size_t) g.N * 3));
	cudaEvent_t start, stop;
	LSF_CUDA(cudaEventCreate(&start));
	LSF_CUDA(cudaEventCreate(&stop));
	int launches = 0;
	LSF_CUDA(cudaEventRecord(start, stream));
	if (stage_ms) {
		// per-stage mode: events around every kernel, accumulated per stage (the total then includes event overhead)
		cudaEvent_t marks[5];
		for (auto& m : marks) LSF_CUDA(cudaEventCreate(&m));
		for (int i = 0; i < 4; i++) stage_ms[i] = 0.0f;
		for (int it = 0; it < iterations; it++) {
			const int n = enqueue_iteration(plan, s, it, false, stream, marks);
			LSF_TRY(n);
			launches += n;
			LSF_CUDA(cudaEventSynchronize(marks[n]));
			for (int i = 0; i < n; i++) {
				float part = 0.0f;
				LSF_CUDA(cudaEventElapsedTime(&part, marks[i], marks[i + 1]));
				stage_ms[i] += part;
			}
		}
		for (auto& m : marks) cudaEventDestroy(m);
	} else {
		for (int it = 0; it < iterations; it++) {
			const int n = enqueue_iteration(plan, s, it, false, stream);
			LSF_TRY(n);
			launches += n;
		}
	}
	finish_deferred(plan, s, iterations, stream);
	if (s.warp_alt != nullptr) launches++;
	LSF_CUDA(cudaEventRecord(stop, stream));
	LSF_CUDA(cudaEventSynchronize(stop));
	LSF_CUDA(cudaGetLastError());
	float ms = 0.0f;
	LSF_CUDA(cudaEventElapsedTime(&ms, start, stop));
	cudaEventDestroy(start);
	cudaEventDestroy(stop);
	if (elapsed_ms) *elapsed_ms = ms;
	if (kernel_launches) *kernel_launches = launches;
	return LSF_OK;
}

// ------------------------------------------------------------------------------------------------ slab decomposition
namespace lsf {
namespace {

// gradient pack of a plane range from a region of the live field (same arithmetic as k_gradient_pack3d)
static __global__ void k_gradient_pack3d_slab(const float* __restrict__ live, int live_planes, int live_origin,
		int X_global, float4* __restrict__ pack, Grid3 pg, int pack_origin) {
	LSF_VOXEL_3D(pg);
	(void) idx;
	if (!in_grid) return;
	const int gx = x + pack_origin;
	if (gx < 0 || gx >= X_global) return;  // stays at the out-of-bounds constants
	const int lx = gx - live_origin;
	const long long YZ = (long long) pg.Y * pg.Z;
	const long long at = lx * YZ + (long long) y * pg.Z + z;
	(void) live_planes;
	float dx = 0.0f;
	if (X_global >= 2) {
		if (gx == 0) dx = live[at + YZ] - live[at];
		else if (gx == X_global - 1) dx = live[at] - live[at - YZ];
		else dx = 0.5f * (live[at + YZ] - live[at - YZ]);
	}
	const float dy = central_difference(live, at, pg.Z, y, pg.Y);
	const float dz = central_difference(live, at, 1, z, pg.Z);
	pack[pg.padded_index(x, y, z)] = make_float4(live[at], dx, dy, dz);
}

// reference downsampleX2_average (resampling.tpp:385-417) between slab allocations: dst plane x <- src planes
// 2x + shift, 2x + shift + 1
template<typename Access>
static __global__ void k_downsample_average3d_slab(Access acc, Grid3 src, Grid3 dst, int dst_begin, int shift) {
	const int z = blockIdx.x * BLOCK_Z + threadIdx.x;
	const int y = blockIdx.y * BLOCK_Y + threadIdx.y;
	const int x = dst_begin + blockIdx.z;
	if (z >= dst.Z || y >= dst.Y) return;
	const int sx = 2 * x + shift, sy = 2 * y, sz = 2 * z;
	auto sum = acc.load(src, sx, sy, sz) + acc.load(src, sx + 1, sy, sz);
	sum = sum + acc.load(src, sx, sy + 1, sz);
	sum = sum + acc.load(src, sx + 1, sy + 1, sz);
	sum = sum + acc.load(src, sx, sy, sz + 1);
	sum = sum + acc.load(src, sx + 1, sy, sz + 1);
	sum = sum + acc.load(src, sx, sy + 1, sz + 1);
	sum = sum + acc.load(src, sx + 1, sy + 1, sz + 1);
	acc.store(dst, x, y, z, sum / 8.0f);
}

static __global__ void k_upsample_nearest3d_slab(const float* __restrict__ src, float* __restrict__ dst, Grid3 sg,
		Grid3 dg, int dst_begin, int dst_origin, int src_origin) {
	const int z = blockIdx.x * BLOCK_Z + threadIdx.x;
	const int y = blockIdx.y * BLOCK_Y + threadIdx.y;
	const int x = dst_begin + blockIdx.z;
	if (z >= dg.Z || y >= dg.Y) return;
	const int sx = ((x + dst_origin) >> 1) - src_origin;
	const long long sidx = ((long long) sx * sg.Y + (y >> 1)) * sg.Z + (z >> 1);
	const long long idx = ((long long) x * dg.Y + y) * dg.Z + z;
	for (int c = 0; c < 3; c++) dst[c * dg.N + idx] = src[c * sg.N + sidx];
}

}  // namespace
}  // namespace lsf

extern "C" int lsf_hier_slab_iteration(const lsf_hier_params* params, const lsf_slab_level* level, int iteration,
		int phase, void* stream_handle) {
	cudaStream_t stream = static_cast<cudaStream_t>(stream_handle);
	LSF_REQUIRE(params && level, "params and level must not be NULL");
	LSF_REQUIRE(phase == 1 || phase == 2, "phase must be 1 or 2");
	LSF_REQUIRE(level->planes > 0 && level->Y > 0 && level->Z > 0 && level->own_begin >= 0
			&& level->own_begin < level->own_end && level->own_end <= level->planes, "invalid slab geometry");
	Plan3 plan;
	plan.tikhonov = params->tikhonov_term_enabled && params->tikhonov_strength > 0.0f;
	plan.use_kernel = params->gradient_kernel_enabled && params->kernel_size > 0 && params->kernel != nullptr;
	if (plan.use_kernel) {
		LSF_TRY(make_taps(params->kernel, params->kernel_size, &plan.taps));
		LSF_REQUIRE(plan.taps.radius >= 1 && plan.taps.radius <= 3, "slab mode supports 3, 5 and 7-tap kernels");
		LSF_REQUIRE(level->own_begin == 0 || level->own_begin >= plan.taps.radius,
				"the low halo (%d planes) is narrower than the kernel radius", level->own_begin);
		LSF_REQUIRE(level->own_end == level->planes || level->planes - level->own_end >= plan.taps.radius,
				"the high halo (%d planes) is narrower than the kernel radius", level->planes - level->own_end);
	}
	plan.rate = params->rate;
	plan.threshold = params->maximum_warp_update_threshold;
	plan.amplifier = params->data_term_amplifier;
	plan.strength = params->tikhonov_strength;
	plan.max_iterations = params->maximum_iteration_count;
	plan.allow_fast_kernels = true;
	plan.stage1_variant = 2;
	plan.lane_xv = plan.tikhonov ? 2 : 4;
	LevelState s;
	s.g = Grid3(level->planes, level->Y, level->Z);
	LSF_REQUIRE(s.g.N * 3 < (1ll << 31), "slab too large for 32-bit voxel indices (%lld voxels): use more ranks", s.g.N);
	s.pack = static_cast<const float4*>(level->pack);
	s.canonical = level->canonical;
	s.warp = level->warp;
	s.g_post = level->g_post;
	s.scratch_a = level->g_pre;
	s.scratch_b = nullptr;
	s.max_sq_bits = level->max_sq_bits;
	s.slab = true;
	s.x_begin = level->own_begin;
	s.x_end = level->own_end;
	s.x_origin = level->x_origin;
	s.X_global = level->X_global;
	s.pack_X = level->pack_planes;
	s.pack_origin = level->pack_origin;
	s.pack_interior_low = level->pack_interior_low;
	s.pack_interior_high = level->pack_interior_high;
	s.violation = level->violation;
	const char* slab_fast = getenv("LSF_SLAB_FAST");
	plan.slab_fast = !(slab_fast && slab_fast[0] == '0');
	Arena arena(stream);
	if (plan.slab_fast && plan.use_kernel && phase == 2) LSF_TRY(arena.alloc(&s.scratch_h, (size_t) s.g.N * 3));
	const int launched = enqueue_iteration(plan, s, iteration, true, stream, nullptr, phase);
	LSF_TRY(launched);
	LSF_CUDA(cudaGetLastError());
	return LSF_OK;
}

extern "C" int lsf_slab_pack_finest(const float* live_region, int live_planes, int live_origin, int X_global, int Y,
		int Z, void* pack, int pack_planes, int pack_origin, void* stream_handle) {
	cudaStream_t stream = static_cast<cudaStream_t>(stream_handle);
	LSF_REQUIRE(live_region && pack && live_planes > 0 && pack_planes > 0 && Y > 0 && Z > 0, "invalid arguments");
	const int lo = std::max(pack_origin, 0), hi = std::min(pack_origin + pack_planes, X_global);
	LSF_REQUIRE(std::max(lo - 1, 0) >= live_origin && std::min(hi + 1, X_global) <= live_origin + live_planes,
			"the live region [%d, %d) does not cover the pack planes [%d, %d) +- 1", live_origin,
			live_origin + live_planes, lo, hi);
	const Grid3 pg(pack_planes, Y, Z);
	const float4 border = make_float4(1.0f, 0.0f, 0.0f, 0.0f);
	k_fill4<<<counted(div_up(pg.padded_count(), 256)), 256, 0, stream>>>(static_cast<float4*>(pack), pg.padded_count(), border);
	k_gradient_pack3d_slab<<<counted(grid3(pg)), block3(), 0, stream>>>(live_region, live_planes, live_origin, X_global,
			static_cast<float4*>(pack), pg, pack_origin);
	LSF_CUDA(cudaGetLastError());
	return LSF_OK;
}

extern "C" int lsf_slab_restrict(int kind, const void* src, int src_planes, int src_origin, int src_Y, int src_Z,
		void* dst, int dst_planes, int dst_origin, int dst_begin, int dst_end, void* stream_handle) {
	cudaStream_t stream = static_cast<cudaStream_t>(stream_handle);
	LSF_REQUIRE(src && dst && src_Y % 2 == 0 && src_Z % 2 == 0, "invalid arguments");
	LSF_REQUIRE(0 <= dst_begin && dst_begin <= dst_end && dst_end <= dst_planes, "invalid destination plane range");
	if (dst_begin == dst_end) return LSF_OK;
	const int shift = 2 * dst_origin - src_origin;
	LSF_REQUIRE(2 * dst_begin + shift >= 0 && 2 * (dst_end - 1) + shift + 1 < src_planes,
			"the source planes do not cover the destination range");
	const Grid3 sg(src_planes, src_Y, src_Z), dg(dst_planes, src_Y / 2, src_Z / 2);
	const dim3 grid(div_up(dg.Z, BLOCK_Z), div_up(dg.Y, BLOCK_Y), dst_end - dst_begin);
	if (kind == 1) {
		const float4 border = make_float4(1.0f, 0.0f, 0.0f, 0.0f);
		if (dst_begin == 0 && dst_end == dst_planes)
			k_fill4<<<counted(div_up(dg.padded_count(), 256)), 256, 0, stream>>>(static_cast<float4*>(dst), dg.padded_count(), border);
		k_downsample_average3d_slab<<<counted(grid), block3(), 0, stream>>>(
				PackAccess { static_cast<const float4*>(src), static_cast<float4*>(dst) }, sg, dg, dst_begin, shift);
	} else {
		k_downsample_average3d_slab<<<counted(grid), block3(), 0, stream>>>(
				PlainAccess { static_cast<const float*>(src), static_cast<float*>(dst) }, sg, dg, dst_begin, shift);
	}
	LSF_CUDA(cudaGetLastError());
	return LSF_OK;
}

extern "C" int lsf_slab_prolong_nearest(const float* src, int src_planes, int src_origin, int src_Y, int src_Z,
		float* dst, int dst_planes, int dst_origin, int dst_begin, int dst_end, void* stream_handle) {
	cudaStream_t stream = static_cast<cudaStream_t>(stream_handle);
	LSF_REQUIRE(src && dst, "invalid arguments");
	LSF_REQUIRE(0 <= dst_begin && dst_begin <= dst_end && dst_end <= dst_planes, "invalid destination plane range");
	if (dst_begin == dst_end) return LSF_OK;
	LSF_REQUIRE(((dst_begin + dst_origin) >> 1) - src_origin >= 0
			&& ((dst_end - 1 + dst_origin) >> 1) - src_origin < src_planes, "the source planes do not cover the range");
	const Grid3 sg(src_planes, src_Y, src_Z), dg(dst_planes, src_Y * 2, src_Z * 2);
	const dim3 grid(div_up(dg.Z, BLOCK_Z), div_up(dg.Y, BLOCK_Y), dst_end - dst_begin);
	k_upsample_nearest3d_slab<<<counted(grid), block3(), 0, stream>>>(src, dst, sg, dg, dst_begin, dst_origin, src_origin);
	LSF_CUDA(cudaGetLastError());
	return LSF_OK;
}
