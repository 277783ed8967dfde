// slavcheva.cu -- host side (C-ABI) of the SobolevFusion / KillingFusion optimizers, the masked re-warp primitive and
// the telemetry statistics, on top of slavcheva.cuh.
//
// Control flow restates reference SobolevOptimizer2d::optimize (cpp/src/nonrigid_optimization/slavcheva/
// sobolev_optimizer2d.cpp:71-119, termination optimizer2d.cpp:76-82) and SlavchevaOptimizer2d.optimize
// (nonrigid_opt/slavcheva/slavcheva_optimizer2d.py:332-408). The termination test runs on the device (k_slav_decide);
// the host polls the status flags once per chunk of iterations.
#include "slavcheva.cuh"
#include "slavcheva_fast.cuh"
#include "kernels3d.cuh"  // k_aos_to_planes / k_planes_to_aos

#include <cmath>
#include <cstring>
#include <algorithm>
#include <vector>

namespace lsf {
namespace {

constexpr int POLL_CHUNK = 8;

inline unsigned blocks_for(long long n) {
	return div_up(n, 256);
}

bool host_finished(const SlavParams& p, int completed, int max_iterations, float max_warp) {
	if (p.semantics == LSF_SEMANTICS_CPP)
		return completed >= p.min_iterations && (completed >= max_iterations || max_warp < p.lower || max_warp > p.upper);
	return !(completed < p.min_iterations || (completed < max_iterations && p.lower < max_warp && max_warp < p.upper));
}

int check_geometry(int nd, const int* dims, SlavGeom* g) {
	LSF_REQUIRE(nd == 2 || nd == 3, "fields must be 2D or 3D, got %d dimensions", nd);
	LSF_REQUIRE(dims != nullptr, "dims is NULL");
	for (int a = 0; a < nd; a++) LSF_REQUIRE(dims[a] >= 2, "every field dimension must be at least 2, got %d", dims[a]);
	*g = make_slav_geom(nd, dims);
	LSF_REQUIRE(g->N * 3 < (1ll << 31), "field too large for 32-bit voxel indices (%lld voxels)", g->N);
	return LSF_OK;
}

template<int D>
int statistics_device(const SlavGeom& g, const float* field, long long component_stride, int voxel_stride,
		const float* canonical, const float* live, float min_threshold, float max_threshold,
		lsf_warp_delta_statistics_t* warp_out, lsf_tsdf_difference_statistics_t* diff_out, Arena& arena,
		cudaStream_t stream) {
	StatsAccumulators* acc;
	LSF_TRY(arena.alloc(&acc, 4));
	StatsAccumulators init[4];
	std::memset(init, 0, sizeof(init));
	for (auto& a : init) a.min_bits = 0x7f800000u;
	LSF_CUDA(cudaMemcpyAsync(acc, init, sizeof(init), cudaMemcpyHostToDevice, stream));
	const unsigned blocks = std::min<unsigned>(blocks_for(g.N), 148 * 8);
	StatsAccumulators host[4];
	auto decode = [&](unsigned long long key, int* location) {
		unsigned order = 0xffffffffu - (unsigned) (key & 0xffffffffull);
		int pos[3] = { 0, 0, 0 };
		for (int a = 0; a < D; a++) {
			pos[a] = (int) (order % (unsigned) g.n[a]);
			order /= (unsigned) g.n[a];
		}
		for (int c = 0; c < 3; c++) location[c] = c < D ? pos[g.comp_axis[c]] : 0;
	};
	auto bits_to_float = [](unsigned bits) {
		float f;
		std::memcpy(&f, &bits, sizeof(f));
		return f;
	};
	if (warp_out) {
		k_warp_statistics<D> <<<counted(blocks), 256, 0, stream>>>(g, field, component_stride, voxel_stride, live, canonical,
				min_threshold, 0, 0.0f, acc);
		LSF_CUDA(cudaMemcpyAsync(host, acc, sizeof(StatsAccumulators), cudaMemcpyDeviceToHost, stream));
		LSF_CUDA(cudaStreamSynchronize(stream));
		const float mean = (float) (host[0].sum / host[0].count);
		k_warp_statistics<D> <<<counted(blocks), 256, 0, stream>>>(g, field, component_stride, voxel_stride, live, canonical,
				min_threshold, 1, mean, acc + 1);
		LSF_CUDA(cudaMemcpyAsync(host + 1, acc + 1, sizeof(StatsAccumulators), cudaMemcpyDeviceToHost, stream));
		LSF_CUDA(cudaStreamSynchronize(stream));
		warp_out->ratio_above_min_threshold = (float) (host[0].above / host[0].count);
		warp_out->length_min = std::sqrt(bits_to_float(host[0].min_bits));
		warp_out->length_max = std::sqrt(bits_to_float((unsigned) (host[0].max_key >> 32)));
		warp_out->length_mean = mean;
		warp_out->length_standard_deviation = (float) std::sqrt(host[1].sum / host[0].count);
		decode(host[0].max_key, warp_out->longest_warp_location);
		warp_out->is_largest_below_min_threshold = warp_out->length_max < min_threshold;
		warp_out->is_largest_above_max_threshold = warp_out->length_max > max_threshold;
	}
	if (diff_out) {
		k_difference_statistics<D> <<<counted(blocks), 256, 0, stream>>>(g, live, canonical, 0, 0.0f, acc + 2);
		LSF_CUDA(cudaMemcpyAsync(host + 2, acc + 2, sizeof(StatsAccumulators), cudaMemcpyDeviceToHost, stream));
		LSF_CUDA(cudaStreamSynchronize(stream));
		const float mean = (float) (host[2].sum / (double) g.N);
		k_difference_statistics<D> <<<counted(blocks), 256, 0, stream>>>(g, live, canonical, 1, mean, acc + 3);
		LSF_CUDA(cudaMemcpyAsync(host + 3, acc + 3, sizeof(StatsAccumulators), cudaMemcpyDeviceToHost, stream));
		LSF_CUDA(cudaStreamSynchronize(stream));
		diff_out->difference_min = bits_to_float(host[2].min_bits);
		diff_out->difference_max = bits_to_float((unsigned) (host[2].max_key >> 32));
		diff_out->difference_mean = mean;
		diff_out->difference_standard_deviation = (float) std::sqrt(host[3].sum / (double) g.N);
		decode(host[2].max_key, diff_out->biggest_difference_location);
	}
	LSF_CUDA(cudaGetLastError());
	return LSF_OK;
}

template<int D>
int optimize_device(const lsf_slavcheva_params* params, const SlavGeom& g, const Taps& taps, bool use_kernel,
		const float* live_in, const float* canonical, float* live_out, float* warp_out_aos, lsf_slavcheva_report* report,
		int collect_statistics, float* max_warps, int max_warps_capacity, lsf_iteration_capture* capture,
		float* capture_dev, Arena& arena, cudaStream_t stream, lsf_warp_delta_statistics_t* iteration_statistics = nullptr,
		int iteration_statistics_capacity = 0, double* iteration_energies = nullptr, int iteration_energies_capacity = 0) {
	const size_t N = (size_t) g.N;
	SlavParams p;
	p.semantics = params->semantics;
	p.data_term_method = params->data_term_method;
	p.smoothing_term_method = params->smoothing_term_method;
	p.level_set = params->level_set_term_enabled && params->semantics != LSF_SEMANTICS_PY_VECTORIZED;
	p.rate = params->gradient_descent_rate;
	p.data_weight = params->data_term_weight;
	p.smoothing_weight = params->smoothing_term_weight;
	p.lambda = params->isomorphic_enforcement_factor;
	p.level_set_weight = params->level_set_term_weight;
	p.lower = params->maximum_warp_length_lower_threshold;
	p.upper = params->maximum_warp_length_upper_threshold;
	p.min_iterations = params->minimum_iteration_count;
	const int max_iterations = params->maximum_iteration_count;
	const int bound = std::max(std::max(max_iterations, p.min_iterations), 0);
	const bool cpp = p.semantics == LSF_SEMANTICS_CPP;
	const bool direct = p.semantics == LSF_SEMANTICS_PY_DIRECT;

	float *live_a, *live_b, *warp, *field_a, *field_b, *field_f;
	unsigned* max_sq_bits;
	int* status;
	LSF_TRY(arena.alloc(&live_a, N));
	LSF_TRY(arena.alloc(&live_b, N));
	LSF_TRY(arena.alloc(&warp, N * D));
	LSF_TRY(arena.alloc(&field_a, N * D));
	LSF_TRY(arena.alloc(&field_b, N * D));
	LSF_TRY(arena.alloc(&field_f, N * D));
	LSF_TRY(arena.alloc(&max_sq_bits, (size_t) bound + 1));
	LSF_TRY(arena.alloc(&status, (size_t) bound + 2));
	LSF_CUDA(cudaMemcpyAsync(live_a, live_in, N * sizeof(float), cudaMemcpyDeviceToDevice, stream));
	LSF_CUDA(cudaMemsetAsync(warp, 0, N * D * sizeof(float), stream));
	LSF_CUDA(cudaMemsetAsync(field_f, 0, N * D * sizeof(float), stream));
	LSF_CUDA(cudaMemsetAsync(max_sq_bits, 0, ((size_t) bound + 1) * sizeof(unsigned), stream));
	LSF_CUDA(cudaMemsetAsync(status, 0, ((size_t) bound + 2) * sizeof(int), stream));
	if (capture) capture->count = 0;

	// reference sobolev_optimizer2d.cpp:77 (initial maximum = upper threshold - 1), slavcheva_optimizer2d.py:339 (inf)
	const float initial_max = cpp ? p.upper - 1.0f : INFINITY;
	int executed = 0;
	float last_max = initial_max;
	std::vector<int> host_status((size_t) bound + 2, 0);
	std::vector<unsigned> host_bits((size_t) bound + 1, 0);
	bool finished = host_finished(p, 0, max_iterations, initial_max) || bound == 0;
	int enqueued = 0;
	trace_point("buffers");
	const unsigned blocks = blocks_for(g.N);
	unsigned char* band_dead = nullptr;
	int *band_list = nullptr, *band_positions = nullptr, *band_counts = nullptr, *leave_list = nullptr, *leave_counts = nullptr;
	const char* legacy_filter = getenv("LSF_SLAV_FAST");  // A/B: LSF_SLAV_FAST=0 keeps the first-generation kernels
	const bool fast_filter = !(legacy_filter && legacy_filter[0] == '0');
	// LSF_SLAV_FUSE_REWARP=1 runs the re-warp in the filter kernel's epilogue. Parity-tested, but measured slower at 256^3
	// (0.64 against 0.56 ms per KillingFusion iteration: the divergent gather inside the marching loop costs more than the
	// 24 B per voxel it saves), so it is off by default.
	const char* band_env = getenv("LSF_SLAV_BAND");  // A/B: LSF_SLAV_BAND=0 keeps the uncompacted four-voxel kernels
	const bool band_compaction = !(band_env && band_env[0] == '0');
	const char* fused_env = getenv("LSF_SLAV_FUSE_REWARP");
	const bool fuse_rewarp = fused_env && fused_env[0] == '1';
	// narrow-band sparse iteration (slavcheva_fast.cuh): LSF_SLAV_SPARSE=0 keeps the dense kernels
	const char* rescan_env = getenv("LSF_SLAV_RESCAN");  // iterations between two scans of the band (default below)
	const int rescan_period = rescan_env && atoi(rescan_env) >= 1 ? atoi(rescan_env) : 32;
	const char* sparse_env = getenv("LSF_SLAV_SPARSE");
	auto aligned_16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
	const bool sparse = D == 3 && cpp && fast_filter && band_compaction && use_kernel && !(sparse_env && sparse_env[0] == '0')
			&& taps.radius >= 1 && taps.radius <= 3 && g.n[0] <= 1024 && g.n[1] <= 1024 && g.n[2] <= 1024 && g.n[2] % 4 == 0
			&& g.N * 3 < (1ll << 31) && aligned_16(live_a)
			&& aligned_16(canonical) && !finished;
	// brick-ordered band with TMA-staged persistent kernels (fifth generation, slavcheva_fast.cuh); LSF_SLAV_BRICK=0 keeps
	// the memory-ordered global list
	const char* brick_env = getenv("LSF_SLAV_BRICK");
	const bool bricked = sparse && !(brick_env && brick_env[0] == '0') && g.n[2] >= SLAV_STAGE_Z;
	const int bricks_x = (g.n[0] + SLAV_BRICK_X - 1) / SLAV_BRICK_X, bricks_y = (g.n[1] + SLAV_BRICK_Y - 1) / SLAV_BRICK_Y,
			bricks_z = (g.n[2] + SLAV_BRICK_Z - 1) / SLAV_BRICK_Z;
	const size_t bricks = (size_t) bricks_x * bricks_y * bricks_z;
	SlavBrickMaps maps;
	unsigned short* brick_list = nullptr;
	int *brick_counts = nullptr, *active_bricks = nullptr, *active_counts = nullptr, *cursors = nullptr;
	if (sparse) {
		const size_t dead_flags = bricked ? bricks : (N + 1023) / 1024;
		LSF_TRY(arena.alloc(&leave_list, N));
		LSF_TRY(arena.alloc(&band_dead, dead_flags));
		LSF_CUDA(cudaMemsetAsync(band_dead, 0, dead_flags, stream));
		if (bricked) {
			// list segments, active-brick list (one count per scan), one work cursor per launch
			const size_t generations = (size_t) bound / rescan_period + 1;
			LSF_TRY(arena.alloc(&brick_list, bricks * SLAV_BRICK_VOXELS));
			LSF_TRY(arena.alloc(&brick_counts, bricks));
			LSF_TRY(arena.alloc(&active_bricks, bricks));
			LSF_TRY(arena.alloc(&active_counts, generations));
			LSF_TRY(arena.alloc(&cursors, ((size_t) bound + 1) * 4));
			LSF_CUDA(cudaMemsetAsync(brick_counts, 0, bricks * sizeof(int), stream));
			LSF_CUDA(cudaMemsetAsync(active_counts, 0, generations * sizeof(int), stream));
			LSF_CUDA(cudaMemsetAsync(cursors, 0, ((size_t) bound + 1) * 4 * sizeof(int), stream));
			// tensor maps of the staged boxes (the three update fields keep their roles in this mode)
			const int R = taps.radius;
			LSF_TRY(make_brick_box_map(&maps.live[0], live_a, 1, (long long) N, g, SLAV_BRICK_X + 2, SLAV_BRICK_Y + 2, SLAV_STAGE_Z));
			LSF_TRY(make_brick_box_map(&maps.live[1], live_b, 1, (long long) N, g, SLAV_BRICK_X + 2, SLAV_BRICK_Y + 2, SLAV_STAGE_Z));
			LSF_TRY(make_brick_box_map(&maps.canonical, canonical, 1, (long long) N, g, SLAV_BRICK_X, SLAV_BRICK_Y, SLAV_BRICK_Z));
			LSF_TRY(make_brick_box_map(&maps.warp, warp, 3, (long long) N, g, SLAV_BRICK_X + 2, SLAV_BRICK_Y + 2, SLAV_STAGE_Z));
			LSF_TRY(make_brick_box_map(&maps.pass[0], field_a, 3, (long long) N, g, SLAV_BRICK_X + 2 * R, SLAV_BRICK_Y, SLAV_BRICK_Z));
			LSF_TRY(make_brick_box_map(&maps.pass[1], field_f, 3, (long long) N, g, SLAV_BRICK_X, SLAV_BRICK_Y + 2 * R, SLAV_BRICK_Z));
			LSF_TRY(make_brick_box_map(&maps.pass[2], field_b, 3, (long long) N, g, SLAV_BRICK_X, SLAV_BRICK_Y, SLAV_STAGE_Z));
		} else {
			LSF_TRY(arena.alloc(&band_list, N));
			LSF_TRY(arena.alloc(&band_positions, N));
		}
		LSF_TRY(arena.alloc(&band_counts, (size_t) bound + 1));
		LSF_TRY(arena.alloc(&leave_counts, (size_t) bound + 1));
		LSF_CUDA(cudaMemsetAsync(band_counts, 0, ((size_t) bound + 1) * sizeof(int), stream));
		LSF_CUDA(cudaMemsetAsync(leave_counts, 0, ((size_t) bound + 1) * sizeof(int), stream));
		// invariants outside the band: update fields zero, both live buffers equal (the warp is zero already)
		LSF_CUDA(cudaMemsetAsync(field_a, 0, N * D * sizeof(float), stream));
		LSF_CUDA(cudaMemsetAsync(field_b, 0, N * D * sizeof(float), stream));
		LSF_CUDA(cudaMemcpyAsync(live_b, live_a, N * sizeof(float), cudaMemcpyDeviceToDevice, stream));
	}
	// small 2D fields: a polling chunk of iterations in one cooperative launch (slavcheva_persistent.cu); LSF_SLAV_PERSISTENT=0
	// keeps one launch per kernel (A/B)
	const char* persistent_env = getenv("LSF_SLAV_PERSISTENT");
	// energy log of the reference's Python optimizer (2D): one kernel per iteration in front of the gradient kernel
	const bool log_energies = D == 2 && !cpp && iteration_energies != nullptr && iteration_energies_capacity > 0;
	double* energies_dev = nullptr;
	if (log_energies) {
		LSF_TRY(arena.alloc(&energies_dev, 3 * (size_t) (bound + 1)));
		LSF_CUDA(cudaMemsetAsync(energies_dev, 0, 3 * (size_t) (bound + 1) * sizeof(double), stream));
	}
	const bool persistent = D == 2 && !sparse && capture_dev == nullptr && !log_energies
			&& !(persistent_env && persistent_env[0] == '0') && g.N <= slav_persistent_capacity();
	// the launch ends by itself at the first finished iteration, so a long chunk costs nothing: the host looks once
	constexpr int PERSISTENT_CHUNK = 128;
	std::vector<SlavIterationCommand> commands(persistent ? PERSISTENT_CHUNK : 0);
	SlavIterationCommand* commands_dev = nullptr;
	if (persistent) LSF_TRY(arena.alloc(&commands_dev, (size_t) PERSISTENT_CHUNK));
	while (!finished) {
		// with per-iteration statistics requested (reference sobolev_optimizer2d.cpp:88-97) the host looks at every iteration
		const int chunk_end = std::min(bound, enqueued + (iteration_statistics ? 1 : (persistent ? PERSISTENT_CHUNK : POLL_CHUNK)));
		for (int it = enqueued; it < chunk_end; it++) {
			SlavGradientArgs ga;
			ga.g = g;
			ga.p = p;
			ga.live = live_a;
			ga.canonical = canonical;
			ga.warp = warp;
			ga.stale = direct ? field_f : nullptr;
			ga.out = field_a;
			ga.status = status;
			ga.iteration = it;
			auto aligned_field = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
			SlavBandArgs band;
			band.g = g;
			band.list = band_list;
			band.positions = band_positions;
			band.dead = band_dead;
			// the band list is rebuilt every `rescan_period` iterations and re-used in between (the band only shrinks)
			const bool rescan = it % rescan_period == 0;
			band.count = band_counts ? band_counts + (it / rescan_period) * rescan_period : nullptr;
			band.leave_list = leave_list;
			band.leave_count = leave_counts ? leave_counts + it : nullptr;
			band.status = status;
			band.iteration = it;
			const unsigned band_blocks = 148 * 8;
			SlavBrickArgs brick;
			brick.list = brick_list;
			brick.brick_count = brick_counts;
			brick.dead = band_dead;
			brick.bricks_y = bricks_y;
			brick.bricks_z = bricks_z;
			brick.leave_list = leave_list;
			brick.leave_count = band.leave_count;
			brick.active = active_bricks;
			brick.active_count = active_counts ? active_counts + it / rescan_period : nullptr;
			brick.cursor = cursors ? cursors + (size_t) it * 4 : nullptr;
			brick.status = status;
			brick.iteration = it;
			if (log_energies) k_slav_energies2d<<<counted(blocks), 256, 0, stream>>>(ga, energies_dev + 3 * (size_t) it);
			const int live_parity = it & 1;  // the live buffers swap once per enqueued iteration
			if (bricked) {
				if (rescan) k_slav_brick_scan<<<counted((unsigned) bricks), 256, 0, stream>>>(ga, brick);
				launch_slav_brick_terms_tma(ga, brick, maps, live_parity, stream);
			} else if (sparse) {
				if (rescan) k_slav_band_scan<<<counted((unsigned) ((g.N + 1023) / 1024)), 256, 0, stream>>>(ga, band);
				k_slav_band_terms<<<counted(band_blocks), 256, 0, stream>>>(ga, band);
			}
			else if (D == 3 && cpp && fast_filter && g.n[2] % 4 == 0 && aligned_field(ga.live) && aligned_field(ga.canonical)
					&& aligned_field(ga.out))
			{
				if (band_compaction) k_slav_gradient_cpp3_band<<<counted((unsigned) ((g.N + 1023) / 1024)), 256, 0, stream>>>(ga);
				else k_slav_gradient_cpp3_v4<<<counted(blocks_for(g.N / 4)), 256, 0, stream>>>(ga);  // four voxels per thread
			}
			else if (persistent) commands[it - enqueued].gradient = ga;
			else
				k_slav_gradient<D> <<<counted(blocks), 256, 0, stream>>>(ga);
			if (persistent) commands[it - enqueued].passes = 0;
			float* final_field = field_a;
			SlavResampleArgs ra;
			ra.g = g;
			ra.p = p;
			ra.live = live_a;
			ra.canonical = canonical;
			ra.warp = warp;
			ra.new_live = live_b;
			ra.band_union_only = cpp ? 1 : 0;  // sobolev_optimizer2d.cpp:134 vs slavcheva_optimizer2d.py:227-228,324-328
			ra.known_values_only = 0;
			ra.substitute_original = 0;
			ra.modify_warp = 1;
			ra.threshold = 1e-6f;
			ra.max_sq_bits = max_sq_bits + it;
			ra.status = status;
			ra.iteration = it;
			bool rewarped = false;  // the re-warp ran in the filter kernel's epilogue
			if (use_kernel) {
				SlavFilterArgs fa;
				fa.g = g;
				fa.original = field_a;
				for (int j = 0; j < LSF_MAX_KERNEL_SIZE; j++) fa.k[j] = taps.k[j];
				fa.size = taps.size;
				fa.radius = taps.radius;
				fa.zero_rule = cpp ? 1 : 2;
				fa.status = status;
				fa.iteration = it;
				// pass order: array axis 0 first (2D: rows then columns, convolution.cpp:69-145; 3D: axes 0, 1, 2)
				const float* in = field_a;
				float* outs[3] = { field_b, field_f, field_b };
				if (bricked) {
					// a -> f (axis 0) -> b (axis 1); the axis-2 pass runs in the re-warp kernel, its result is not stored
					ra.update = nullptr;
					ra.gradient_field = nullptr;
					fa.in = field_a;
					fa.out = field_f;
					fa.axis = 0;
					brick.cursor++;
					if (taps.radius == 1) launch_slav_brick_filter_tma<1, 0>(fa, brick, maps.pass[0], stream);
					else if (taps.radius == 2) launch_slav_brick_filter_tma<2, 0>(fa, brick, maps.pass[0], stream);
					else launch_slav_brick_filter_tma<3, 0>(fa, brick, maps.pass[0], stream);
					fa.in = field_f;
					fa.out = field_b;
					fa.axis = 1;
					brick.cursor++;
					if (taps.radius == 1) launch_slav_brick_filter_tma<1, 1>(fa, brick, maps.pass[1], stream);
					else if (taps.radius == 2) launch_slav_brick_filter_tma<2, 1>(fa, brick, maps.pass[1], stream);
					else launch_slav_brick_filter_tma<3, 1>(fa, brick, maps.pass[1], stream);
					fa.in = field_b;
					fa.out = nullptr;
					fa.axis = 2;
					brick.cursor++;
					if (taps.radius == 1) launch_slav_brick_filter_resample_tma<1>(fa, ra, brick, maps, live_parity, stream);
					else if (taps.radius == 2) launch_slav_brick_filter_resample_tma<2>(fa, ra, brick, maps, live_parity, stream);
					else launch_slav_brick_filter_resample_tma<3>(fa, ra, brick, maps, live_parity, stream);
					k_slav_brick_leave_decide<<<counted(64u), 256, 0, stream>>>(brick, g.N, live_b, live_a, field_a, field_b, field_f, p,
							max_sq_bits, status, max_iterations);
					rewarped = true;
					outs[D - 1] = field_f;  // no buffer rotation: the three update fields keep their roles
				} else if (sparse) {
					// a -> f (axis 0) -> b (axis 1) -> a (axis 2), at the band voxels only
					float* band_outs[3] = { field_f, field_b, field_a };
					for (int axis = 0; axis < D; axis++) {
						fa.in = in;
						fa.out = band_outs[axis];
						fa.axis = axis;
						if (taps.radius == 1) k_slav_band_filter_axis<1> <<<counted(band_blocks), 256, 0, stream>>>(fa, band);
						else if (taps.radius == 2) k_slav_band_filter_axis<2> <<<counted(band_blocks), 256, 0, stream>>>(fa, band);
						else k_slav_band_filter_axis<3> <<<counted(band_blocks), 256, 0, stream>>>(fa, band);
						in = band_outs[axis];
					}
					outs[D - 1] = field_a;
				} else if (D == 3 && cpp && fast_filter && slav_fast_filter_supported(g, taps, field_a, field_b)) {
					// second generation: axis-0 marching kernel, then axes 1 and 2 in one kernel (slavcheva_fast.cuh) whose
					// epilogue also re-warps the live field (the filtered field itself is then never stored)
					ra.update = nullptr;
					ra.gradient_field = nullptr;
					const SlavResampleArgs* fused = fuse_rewarp ? &ra : nullptr;
					if (taps.radius == 1) launch_slav_fast_filter<1>(g, taps, field_a, field_f, field_b, status, it, stream, fused);
					else if (taps.radius == 2) launch_slav_fast_filter<2>(g, taps, field_a, field_f, field_b, status, it, stream, fused);
					else launch_slav_fast_filter<3>(g, taps, field_a, field_f, field_b, status, it, stream, fused);
					rewarped = fuse_rewarp;
				} else {
					for (int axis = 0; axis < D; axis++) {
						fa.in = in;
						fa.out = outs[axis];
						fa.axis = axis;
						if (persistent) commands[it - enqueued].pass[commands[it - enqueued].passes++] = fa;
						else k_slav_filter_axis<D> <<<counted(blocks), 256, 0, stream>>>(fa);
						in = outs[axis];
					}
				}
				final_field = outs[D - 1];
			}
			ra.update = final_field;
			ra.gradient_field = direct ? final_field : nullptr;
			auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
			if (rewarped) {
				// done
			} else if (sparse) {
				k_slav_band_resample<<<counted(band_blocks), 256, 0, stream>>>(ra, band);
				k_slav_band_leave<<<counted(64u), 256, 0, stream>>>(band, live_b, live_a, field_a, field_b, field_f);
			} else if (D == 3 && fast_filter && g.n[2] % 4 == 0 && aligned16(ra.update) && aligned16(ra.live)
					&& aligned16(ra.canonical) && aligned16(ra.warp) && aligned16(ra.new_live))
			{
				if (band_compaction && cpp && ra.band_union_only && !ra.known_values_only)
					k_slav_resample_band<D> <<<counted((unsigned) ((g.N + 1023) / 1024)), 256, 0, stream>>>(ra);
				else k_slav_resample_v4<D> <<<counted(blocks_for(g.N / 4)), 256, 0, stream>>>(ra);  // four voxels per thread
			}
			else if (persistent) commands[it - enqueued].resample = ra;
			else
				k_slav_resample<D> <<<counted(blocks), 256, 0, stream>>>(ra);
			if (!bricked && !persistent) k_slav_decide<<<counted(1u), 1, 0, stream>>>(p, max_sq_bits, status, it, max_iterations);
			// the filtered field becomes the persistent gradient field read as `stale` next iteration; the live buffers
			// swap. Both swaps also happen for iterations the device skips (status set): skipped kernels write nothing,
			// and the results are read from the buffers of the last executed iteration (tracked below).
			if (final_field != field_f) {
				if (final_field == field_a) std::swap(field_a, field_f);
				else std::swap(field_b, field_f);
			}
			std::swap(live_a, live_b);
			if (capture_dev != nullptr && it < capture->max_iterations)
				k_planes_to_aos<<<counted(blocks), 256, 0, stream>>>(warp, capture_dev + (size_t) it * N * D, g.N, D);
		}
		if (persistent) {
			// pageable source: the driver has copied it out when the call returns, `commands` is free for the next chunk
			LSF_CUDA(cudaMemcpyAsync(commands_dev, commands.data(), (size_t) (chunk_end - enqueued) * sizeof(SlavIterationCommand),
					cudaMemcpyHostToDevice, stream));
			SlavOptimizerBuffers buffers = { { warp, field_a, field_b, field_f }, { live_a, live_b, const_cast<float*>(canonical) },
					g.n[0], g.n[1], use_kernel ? taps.radius : 0 };
			LSF_TRY(launch_slav_persistent2d(commands_dev, chunk_end - enqueued, p, g.N, max_sq_bits, status, enqueued, max_iterations,
					stream, D == 2 ? &buffers : nullptr));
		}
		LSF_CUDA(cudaGetLastError());
		trace_point("chunk enqueued");
		LSF_CUDA(cudaMemcpyAsync(host_status.data() + enqueued + 1, status + enqueued + 1,
				(size_t) (chunk_end - enqueued) * sizeof(int), cudaMemcpyDeviceToHost, stream));
		LSF_CUDA(cudaMemcpyAsync(host_bits.data() + enqueued, max_sq_bits + enqueued,
				(size_t) (chunk_end - enqueued) * sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
		LSF_CUDA(cudaStreamSynchronize(stream));
		trace_point("chunk waited");
		for (int it = enqueued; it < chunk_end && !finished; it++) {
			float sq;
			std::memcpy(&sq, &host_bits[it], sizeof(float));
			last_max = std::sqrt(sq);
			if (max_warps && it < max_warps_capacity) max_warps[it] = last_max;
			executed = it + 1;
			if (host_status[it + 1]) finished = true;
		}
		if (iteration_statistics && executed == chunk_end && executed <= iteration_statistics_capacity) {
			// statistics of the warp field over the band union of (canonical, warped live) after this iteration
			LSF_TRY(statistics_device<D>(g, warp, g.N, 1, canonical, live_a, p.lower, p.upper, &iteration_statistics[executed - 1],
					nullptr, arena, stream));
		}
		// live buffers were swapped once per ENQUEUED iteration; undo the swaps of the skipped ones
		if (finished && ((chunk_end - executed) & 1)) std::swap(live_a, live_b);
		enqueued = chunk_end;
		if (enqueued >= bound) finished = true;
	}
	// results: live_a holds the live field after the last executed iteration, `warp` its warp field
	LSF_CUDA(cudaMemcpyAsync(live_out, live_a, N * sizeof(float), cudaMemcpyDeviceToDevice, stream));
	if (warp_out_aos) k_planes_to_aos<<<counted(blocks), 256, 0, stream>>>(warp, warp_out_aos, g.N, D);
	if (capture) capture->count = capture_dev ? std::min(executed, capture->max_iterations) : 0;
	if (iteration_energies != nullptr && iteration_energies_capacity > 0) {
		const int rows = std::min(executed, iteration_energies_capacity);
		std::memset(iteration_energies, 0, 3 * (size_t) iteration_energies_capacity * sizeof(double));
		if (log_energies && rows > 0) {
			LSF_CUDA(cudaMemcpyAsync(iteration_energies, energies_dev, 3 * (size_t) rows * sizeof(double), cudaMemcpyDeviceToHost,
					stream));
			LSF_CUDA(cudaStreamSynchronize(stream));
		}
	}
	if (report) {
		std::memset(report, 0, sizeof(*report));
		report->iteration_count = executed;
		report->iteration_limit_reached = executed >= max_iterations;
		report->last_max_warp_length = last_max;
		if (collect_statistics) {
			// reference sobolev_optimizer2d.cpp:98-117: statistics of the last warp field over the band union of
			// (canonical, warped live), difference statistics of (canonical, warped live)
			LSF_TRY(statistics_device<D>(g, warp, g.N, 1, canonical, live_a, p.lower, p.upper, &report->warp_delta_statistics,
					&report->tsdf_difference_statistics, arena, stream));
			report->has_statistics = 1;
		}
	}
	LSF_CUDA(cudaGetLastError());
	return LSF_OK;
}

}  // namespace

int statistics_on_device(int nd, const int* dims, const float* field, long long component_stride, int voxel_stride,
		const float* canonical, const float* live, float min_threshold, float max_threshold,
		lsf_warp_delta_statistics_t* warp_out, lsf_tsdf_difference_statistics_t* diff_out, Arena& arena,
		cudaStream_t stream) {
	SlavGeom g;
	LSF_TRY(check_geometry(nd, dims, &g));
	if (nd == 2)
		return statistics_device<2>(g, field, component_stride, voxel_stride, canonical, live, min_threshold,
				max_threshold, warp_out, diff_out, arena, stream);
	return statistics_device<3>(g, field, component_stride, voxel_stride, canonical, live, min_threshold, max_threshold,
			warp_out, diff_out, arena, stream);
}

}  // namespace lsf

using namespace lsf;

extern "C" int lsf_slavcheva_optimize_logged(const lsf_slavcheva_params* params, const float* live, const float* canonical,
		int nd, const int* dims, float* live_out, float* warp_out, int memory_kind, lsf_slavcheva_report* report,
		int collect_statistics, float* max_warps, int max_warps_capacity, lsf_iteration_capture* capture,
		lsf_warp_delta_statistics_t* iteration_statistics, int iteration_statistics_capacity, double* iteration_energies,
		int iteration_energies_capacity, void* stream_handle) {
	cudaStream_t stream = static_cast<cudaStream_t>(stream_handle);
	LSF_REQUIRE(params != nullptr, "params is NULL");
	LSF_REQUIRE(live && canonical && live_out, "live, canonical and live_out must not be NULL");
	SlavGeom g;
	LSF_TRY(check_geometry(nd, dims, &g));
	// reference slavcheva_optimizer2d.py:157-161 (ValueError) and the square-only index decoding of the C++ code
	// (field_warping.cpp:80-82): 2D fields must be square
	LSF_REQUIRE(nd == 3 || dims[0] == dims[1], "2D fields must be square, got %d x %d", dims[0], dims[1]);
	LSF_REQUIRE(params->semantics == LSF_SEMANTICS_CPP || nd == 2,
			"the Python semantics exist for 2D fields only (the reference has no 3D slavcheva optimizer)");
	LSF_REQUIRE(params->semantics >= 0 && params->semantics <= 2, "unknown semantics %d", params->semantics);
	Taps taps;
	std::memset(&taps, 0, sizeof(taps));
	const bool use_kernel = params->sobolev_smoothing_enabled && params->sobolev_kernel && params->sobolev_kernel_size > 0;
	if (use_kernel) LSF_TRY(make_taps(params->sobolev_kernel, params->sobolev_kernel_size, &taps));
	const size_t N = (size_t) g.N;
	trace_point("begin");
	Arena arena(stream);
	const float *live_dev, *canonical_dev;
	LSF_TRY(to_device(arena, live, N, memory_kind, stream, &live_dev));
	LSF_TRY(to_device(arena, canonical, N, memory_kind, stream, &canonical_dev));
	trace_point("inputs");
	float *live_out_dev = live_out, *warp_out_dev = warp_out, *capture_dev = nullptr;
	if (memory_kind == LSF_HOST) {
		LSF_TRY(arena.alloc(&live_out_dev, N));
		if (warp_out) LSF_TRY(arena.alloc(&warp_out_dev, N * nd));
	}
	const bool capturing = capture && capture->max_iterations > 0 && capture->buffer;
	if (capturing) {
		if (memory_kind == LSF_HOST) LSF_TRY(arena.alloc(&capture_dev, (size_t) capture->max_iterations * N * nd));
		else capture_dev = capture->buffer;
	}
	if (nd == 2)
		LSF_TRY(optimize_device<2>(params, g, taps, use_kernel, live_dev, canonical_dev, live_out_dev, warp_out_dev, report,
				collect_statistics, max_warps, max_warps_capacity, capture, capture_dev, arena, stream, iteration_statistics,
				iteration_statistics_capacity, iteration_energies, iteration_energies_capacity));
	else
		LSF_TRY(optimize_device<3>(params, g, taps, use_kernel, live_dev, canonical_dev, live_out_dev, warp_out_dev, report,
				collect_statistics, max_warps, max_warps_capacity, capture, capture_dev, arena, stream, iteration_statistics,
				iteration_statistics_capacity));
	trace_point("optimized");
	if (memory_kind == LSF_HOST) {
		if (capture_dev && capture->count > 0)
			LSF_CUDA(cudaMemcpyAsync(capture->buffer, capture_dev, (size_t) capture->count * N * nd * sizeof(float),
					cudaMemcpyDeviceToHost, stream));
		if (warp_out)
			LSF_CUDA(cudaMemcpyAsync(warp_out, warp_out_dev, N * nd * sizeof(float), cudaMemcpyDeviceToHost, stream));
		LSF_TRY(from_device(live_out_dev, live_out, N, LSF_HOST, stream));
	}
	trace_point("result");
	trace_point("end");
	return LSF_OK;
}

extern "C" int lsf_slavcheva_optimize(const lsf_slavcheva_params* params, const float* live, const float* canonical,
		int nd, const int* dims, float* live_out, float* warp_out, int memory_kind, lsf_slavcheva_report* report,
		int collect_statistics, float* max_warps, int max_warps_capacity, lsf_iteration_capture* capture,
		void* stream_handle) {
	return lsf_slavcheva_optimize_logged(params, live, canonical, nd, dims, live_out, warp_out, memory_kind, report,
			collect_statistics, max_warps, max_warps_capacity, capture, nullptr, 0, nullptr, 0, stream_handle);
}

extern "C" int lsf_warp_advanced(const float* live, const float* canonical, float* warp, int nd, const int* dims,
		int band_union_only, int known_values_only, int substitute_original, float truncation_float_threshold,
		int modify_warp, float* live_out, int memory_kind, void* stream_handle) {
	cudaStream_t stream = static_cast<cudaStream_t>(stream_handle);
	LSF_REQUIRE(live && canonical && warp && live_out, "live, canonical, warp and live_out must not be NULL");
	SlavGeom g;
	LSF_TRY(check_geometry(nd, dims, &g));
	LSF_REQUIRE(nd == 3 || dims[0] == dims[1], "2D fields must be square, got %d x %d", dims[0], dims[1]);
	const size_t N = (size_t) g.N;
	Arena arena(stream);
	const float *live_dev, *canonical_dev, *warp_in_dev;
	LSF_TRY(to_device(arena, live, N, memory_kind, stream, &live_dev));
	LSF_TRY(to_device(arena, canonical, N, memory_kind, stream, &canonical_dev));
	LSF_TRY(to_device(arena, warp, N * nd, memory_kind, stream, &warp_in_dev));
	float *planes_in, *planes_out, *live_out_dev = live_out;
	LSF_TRY(arena.alloc(&planes_in, N * nd));
	LSF_TRY(arena.alloc(&planes_out, N * nd));
	if (memory_kind == LSF_HOST || live_out == live) LSF_TRY(arena.alloc(&live_out_dev, N));
	const unsigned blocks = blocks_for(g.N);
	k_aos_to_planes<<<counted(blocks), 256, 0, stream>>>(warp_in_dev, planes_in, g.N, nd);
	SlavResampleArgs ra;
	std::memset(&ra, 0, sizeof(ra));
	ra.g = g;
	ra.p.semantics = LSF_SEMANTICS_CPP;
	ra.live = live_dev;
	ra.canonical = canonical_dev;
	ra.update = planes_in;
	ra.gradient_field = nullptr;
	ra.warp = planes_out;
	ra.new_live = live_out_dev;
	ra.band_union_only = band_union_only;
	ra.known_values_only = known_values_only;
	ra.substitute_original = substitute_original;
	ra.modify_warp = modify_warp;
	ra.threshold = truncation_float_threshold;
	ra.max_sq_bits = nullptr;
	ra.status = nullptr;
	ra.iteration = 0;
	if (nd == 2) k_slav_resample<2> <<<counted(blocks), 256, 0, stream>>>(ra);
	else k_slav_resample<3> <<<counted(blocks), 256, 0, stream>>>(ra);
	LSF_CUDA(cudaGetLastError());
	if (modify_warp) {
		float* warp_dev = const_cast<float*>(warp_in_dev);  // staged copy (host) or the caller's buffer (device)
		k_planes_to_aos<<<counted(blocks), 256, 0, stream>>>(planes_out, warp_dev, g.N, nd);
		if (memory_kind == LSF_HOST)
			LSF_CUDA(cudaMemcpyAsync(warp, warp_dev, N * nd * sizeof(float), cudaMemcpyDeviceToHost, stream));
	}
	if (memory_kind == LSF_HOST) return from_device(live_out_dev, live_out, N, LSF_HOST, stream);
	if (live_out_dev != live_out)
		LSF_CUDA(cudaMemcpyAsync(live_out, live_out_dev, N * sizeof(float), cudaMemcpyDeviceToDevice, stream));
	return LSF_OK;
}

extern "C" int lsf_warp_delta_statistics(const float* warp, const float* canonical, const float* live, int nd,
		const int* dims, float min_threshold, float max_threshold, lsf_warp_delta_statistics_t* out, int memory_kind,
		void* stream_handle) {
	cudaStream_t stream = static_cast<cudaStream_t>(stream_handle);
	LSF_REQUIRE(warp && canonical && live && out, "warp, canonical, live and out must not be NULL");
	SlavGeom g;
	LSF_TRY(check_geometry(nd, dims, &g));
	Arena arena(stream);
	const float *warp_dev, *canonical_dev, *live_dev;
	LSF_TRY(to_device(arena, warp, (size_t) g.N * nd, memory_kind, stream, &warp_dev));
	LSF_TRY(to_device(arena, canonical, (size_t) g.N, memory_kind, stream, &canonical_dev));
	LSF_TRY(to_device(arena, live, (size_t) g.N, memory_kind, stream, &live_dev));
	if (nd == 2)
		return statistics_device<2>(g, warp_dev, 1, 2, canonical_dev, live_dev, min_threshold, max_threshold, out, nullptr,
				arena, stream);
	return statistics_device<3>(g, warp_dev, 1, 3, canonical_dev, live_dev, min_threshold, max_threshold, out, nullptr,
			arena, stream);
}

extern "C" int lsf_tsdf_difference_statistics(const float* canonical, const float* live, int nd, const int* dims,
		lsf_tsdf_difference_statistics_t* out, int memory_kind, void* stream_handle) {
	cudaStream_t stream = static_cast<cudaStream_t>(stream_handle);
	LSF_REQUIRE(canonical && live && out, "canonical, live and out must not be NULL");
	SlavGeom g;
	LSF_TRY(check_geometry(nd, dims, &g));
	Arena arena(stream);
	const float *canonical_dev, *live_dev;
	LSF_TRY(to_device(arena, canonical, (size_t) g.N, memory_kind, stream, &canonical_dev));
	LSF_TRY(to_device(arena, live, (size_t) g.N, memory_kind, stream, &live_dev));
	if (nd == 2)
		return statistics_device<2>(g, nullptr, 0, 0, canonical_dev, live_dev, 0.0f, 0.0f, nullptr, out, arena, stream);
	return statistics_device<3>(g, nullptr, 0, 0, canonical_dev, live_dev, 0.0f, 0.0f, nullptr, out, arena, stream);
}
