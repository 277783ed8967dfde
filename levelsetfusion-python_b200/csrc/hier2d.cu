// hier2d.cu -- 2D hierarchical optimizer driver: the reference's Optimizer<MatrixXf,MatrixXv2f>
// (cpp/src/nonrigid_optimization/hierarchical/optimizer.tpp:83-212; Python twin
// nonrigid_opt/hierarchical/hierarchical_optimizer2d.py:123-248). Same device-side termination scheme as
// hier3d.cu.
#include "kernels2d.cuh"
#include "slavcheva.cuh"  // statistics_on_device
#include "hier_telemetry.cuh"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <algorithm>

namespace lsf {

namespace {

constexpr int POLL_CHUNK = 16;

struct Plan2 {
	bool tikhonov = false, use_kernel = false, linear = false;
	int level_count = 0;
	Taps taps;
	float rate = 0, threshold = 0, amplifier = 0, strength = 0;
	int max_iterations = 0;
	Grid2 level_grid[LSF_MAX_LEVELS];
};

int make_plan(const lsf_hier_params* p, int H, int W, Plan2* plan) {
	LSF_REQUIRE(p != nullptr, "params is NULL");
	LSF_REQUIRE(H > 0 && W > 0, "field dimensions must be positive, got %d x %d", H, W);
	LSF_REQUIRE(is_power_of_two(p->maximum_chunk_size),
			"The argument 'maximum_chunk_size' must be an integer power of 2, i.e. 4, 8, 16, etc.");
	const int power = (int) std::log2((double) p->maximum_chunk_size);
	const int max_level_count = (int) std::min(std::log2((double) H), std::log2((double) W)) + 1;
	LSF_REQUIRE(max_level_count > power, "Maximum chunk size too large for the field size.");
	plan->level_count = power + 1;
	LSF_REQUIRE(plan->level_count <= LSF_MAX_LEVELS, "too many pyramid levels (%d)", plan->level_count);
	LSF_REQUIRE(p->resampling_strategy == LSF_RESAMPLING_LINEAR
			|| p->resampling_strategy == LSF_RESAMPLING_NEAREST_AND_AVERAGE, "Unknown resampling strategy %d",
			p->resampling_strategy);
	plan->linear = p->resampling_strategy == LSF_RESAMPLING_LINEAR;
	Grid2 g(H, W);
	for (int level = plan->level_count - 1; level >= 0; level--) {
		plan->level_grid[level] = g;
		if (level > 0) {
			if (plan->linear) {
				// reference resampling.tpp:424-425
				LSF_REQUIRE(g.H % 2 == 0 && g.W % 2 == 0 && g.H > 2 && g.W > 2,
						"Each dimension of the argument 'field' must be divisible by 2 and greater than 2.");
			} else {
				// reference resampling.tpp:361-362
				LSF_REQUIRE(is_power_of_two(g.H) && is_power_of_two(g.W),
						"The argument 'field' must have a power of two for each dimension.");
			}
		}
		g = g.half();
	}
	plan->tikhonov = p->tikhonov_term_enabled && p->tikhonov_strength > 0.0f;
	plan->use_kernel = p->gradient_kernel_enabled && p->kernel_size > 0 && p->kernel != nullptr;
	if (plan->use_kernel) LSF_TRY(make_taps(p->kernel, p->kernel_size, &plan->taps));
	plan->rate = p->rate;
	plan->threshold = p->maximum_warp_update_threshold;
	plan->amplifier = p->data_term_amplifier;
	plan->strength = p->tikhonov_strength;
	plan->max_iterations = p->maximum_iteration_count;
	return LSF_OK;
}

struct LevelState2 {
	Grid2 g;
	const float4* pack = nullptr;
	const float* canonical = nullptr;
	float* warp = nullptr;
	float* g_post = nullptr;
	float* scratch_a = nullptr;
	unsigned* max_sq_bits = nullptr;
};

// iterations [first, first + count) of a level in one cooperative launch (small fields)
int enqueue_level_persistent(const Plan2& plan, LevelState2& s, int first, int count, cudaStream_t stream) {
	HierIterArgs2 a;
	a.pack = s.pack;
	a.canonical = s.canonical;
	a.warp = s.warp;
	a.warp_out = s.warp;
	a.g_prev = s.g_post;
	a.g_out = nullptr;
	a.g = s.g;
	a.amplifier = plan.amplifier;
	a.strength = plan.strength;
	a.rate = plan.rate;
	a.threshold = plan.threshold;
	a.max_sq_bits = s.max_sq_bits;
	a.iteration = first;
	a.check_convergence = 1;
	ConvArgs2 c;
	c.in = nullptr;
	c.out = nullptr;
	c.warp = s.warp;
	c.g = s.g;
	c.taps = plan.taps;
	c.rate = plan.rate;
	c.threshold = plan.threshold;
	c.max_sq_bits = s.max_sq_bits;
	c.iteration = first;
	c.check_convergence = 1;
	c.channels = 2;
	c.preserve_zeros = 0;
	return launch_hier2d_persistent(a, c, plan.tikhonov, plan.use_kernel, s.g_post, s.scratch_a, first, count, stream);
}

void enqueue_iteration(const Plan2& plan, LevelState2& s, int iteration, cudaStream_t stream) {
	HierIterArgs2 a;
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
	a.check_convergence = 1;
	const dim3 grid = grid2(s.g), block = block3();
	if (!plan.use_kernel) {
		if (plan.tikhonov) {
			a.g_out = s.scratch_a;
			k_hier_gradient2d<true, true> <<<counted(grid), block, 0, stream>>>(a);
			std::swap(s.g_post, s.scratch_a);
		} else {
			a.g_out = nullptr;
			k_hier_gradient2d<false, true> <<<counted(grid), block, 0, stream>>>(a);
		}
		return;
	}
	// g_pre -> g_post (stage 1), rows pass g_post -> scratch_a, columns pass scratch_a -> g_post (+ update)
	a.g_out = s.scratch_a;
	if (plan.tikhonov) k_hier_gradient2d<true, false> <<<counted(grid), block, 0, stream>>>(a);
	else k_hier_gradient2d<false, false> <<<counted(grid), block, 0, stream>>>(a);
	ConvArgs2 c;
	c.g = s.g;
	c.taps = plan.taps;
	c.rate = plan.rate;
	c.threshold = plan.threshold;
	c.max_sq_bits = s.max_sq_bits;
	c.iteration = iteration;
	c.check_convergence = 1;
	c.channels = 2;
	c.preserve_zeros = 0;
	c.warp = s.warp;
	// stage 1 wrote scratch_a and no longer needs g_post (g_prev): reuse it as the rows-pass output
	c.in = s.scratch_a;
	c.out = s.g_post;
	k_convolve_axis2d<0, false> <<<counted(grid), block, 0, stream>>>(c);
	c.in = s.g_post;
	c.out = s.scratch_a;
	k_convolve_axis2d<1, true> <<<counted(grid), block, 0, stream>>>(c);
	std::swap(s.g_post, s.scratch_a);
}

static __global__ void k_unpack_live2d(const float4* __restrict__ pack, float* __restrict__ out, Grid2 g) {
	LSF_PIXEL_2D(g);
	if (!in_grid) return;
	out[idx] = pack[g.padded_index(row, col)].x;
}

// diff = warped live - canonical and the data-term gradient of the current warp (reference optimizer.tpp:186-194)
static __global__ void k_telemetry_terms2d(const float4* __restrict__ pack, const float* __restrict__ canonical,
		const float* __restrict__ warp, float* __restrict__ diff, float* __restrict__ data_planes, Grid2 g) {
	LSF_PIXEL_2D(g);
	if (!in_grid) return;
	const float4 s = gather4_2d(pack, g, row, col, warp[idx], warp[g.N + idx]);
	const float d = s.x - canonical[idx];
	diff[idx] = d;
	data_planes[idx] = s.y * d;
	data_planes[g.N + idx] = s.z * d;
}

// One level with per-iteration telemetry (see lsf_iteration_sink in include/lsf_b200.h and hier3d.cu)
int run_level_with_telemetry(const Plan2& plan, LevelState2& s, int level, const lsf_iteration_sink* sink,
		TelemetryScratch& t, cudaStream_t stream, int* executed_out, float* last_max_out) {
	const long long N = s.g.N;
	const int dims[3] = { s.g.H, s.g.W, 1 };
	lsf_iteration_record record;
	std::memset(&record, 0, sizeof(record));
	record.level = level;
	for (int i = 0; i < 3; i++) record.dims[i] = dims[i];
	k_unpack_live2d<<<counted(grid2(s.g)), block3(), 0, stream>>>(s.pack, t.live_level, s.g);
	if (level == 0 && sink->want_fields) {
		LSF_CUDA(cudaMemsetAsync(t.data_planes, 0, (size_t) N * 2 * sizeof(float), stream));
		if (plan.tikhonov) LSF_CUDA(cudaMemsetAsync(t.tikhonov_planes, 0, (size_t) N * 2 * sizeof(float), stream));
		record.iteration = -1;
		LSF_TRY(telemetry_fields(t, N, 2, t.data_planes, plan.tikhonov, &record, stream));
		sink->callback(sink->user, &record);
	}
	int executed = 0;
	float last_max = FLT_MAX;
	unsigned bits = 0;
	for (int it = 0; it < plan.max_iterations; it++) {
		if (last_max < plan.threshold) break;  // reference optimizer.tpp:149,166-171
		k_telemetry_terms2d<<<counted(grid2(s.g)), block3(), 0, stream>>>(s.pack, s.canonical, s.warp, t.diff, t.data_planes, s.g);
		if (plan.tikhonov)
			k_laplacian_planes2d<<<counted(grid2(s.g)), block3(), 0, stream>>>(s.g_post, t.tikhonov_planes, 2, s.g);
		record.iteration = it;
		if (sink->want_statistics) LSF_TRY(telemetry_statistics(t, N, 2, dims, s.g_post, plan.tikhonov, &record, stream));
		enqueue_iteration(plan, s, it, stream);
		LSF_CUDA(cudaMemcpyAsync(&bits, s.max_sq_bits + it, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
		LSF_CUDA(cudaStreamSynchronize(stream));
		float sq;
		std::memcpy(&sq, &bits, sizeof(float));
		last_max = std::sqrt(sq);
		executed = it + 1;
		record.max_update_length = last_max;
		if (sink->want_fields) LSF_TRY(telemetry_fields(t, N, 2, s.warp, plan.tikhonov, &record, stream));
		sink->callback(sink->user, &record);
	}
	*executed_out = executed;
	*last_max_out = last_max;
	return LSF_OK;
}

int hier_optimize_2d(const lsf_hier_params* params, const float* canonical, const float* live, int H,
		int W, float* warp_out, int memory_kind, lsf_level_report* reports, int collect_reports,
		lsf_iteration_capture* capture, const lsf_iteration_sink* sink, void* stream_handle) {
	cudaStream_t stream = static_cast<cudaStream_t>(stream_handle);
	Plan2 plan;
	LSF_TRY(make_plan(params, H, W, &plan));
	LSF_REQUIRE(canonical && live && warp_out, "canonical, live and warp_out must not be NULL");
	trace_point("begin");
	const int L = plan.level_count;
	const Grid2 finest = plan.level_grid[L - 1];
	const size_t N = (size_t) finest.N;
	Arena arena(stream);
	const float *canonical_dev, *live_dev;
	LSF_TRY(to_device(arena, canonical, N, memory_kind, stream, &canonical_dev));
	LSF_TRY(to_device(arena, live, N, memory_kind, stream, &live_dev));
	float* out_dev = warp_out;
	if (memory_kind == LSF_HOST) LSF_TRY(arena.alloc(&out_dev, N * 2));
	trace_point("inputs");

	// pyramids (reference pyramid.tpp:51-74): live + its full-resolution gradient restricted together
	std::vector<float4*> packs(L, nullptr);
	std::vector<const float*> canonicals(L, nullptr);
	const float4 border = make_float4(1.0f, 0.0f, 0.0f, 0.0f);
	for (int level = L - 1; level >= 0; level--) {
		const Grid2& g = plan.level_grid[level];
		LSF_TRY(arena.alloc(&packs[level], (size_t) g.padded_count()));
		k_fill4<<<counted(div_up(g.padded_count(), 256)), 256, 0, stream>>>(packs[level], g.padded_count(), border);
		if (level == L - 1) {
			k_gradient_pack2d<<<counted(grid2(g)), block3(), 0, stream>>>(live_dev, packs[level], g);
			canonicals[level] = canonical_dev;
		} else {
			const Grid2& src = plan.level_grid[level + 1];
			float* canonical_level = nullptr;
			LSF_TRY(arena.alloc(&canonical_level, (size_t) g.N));
			if (plan.linear) {
				k_downsample_linear2d<<<counted(grid2(g)), block3(), 0, stream>>>(PackAccess2 { packs[level + 1], packs[level] },
						src, g);
				k_downsample_linear2d<<<counted(grid2(g)), block3(), 0, stream>>>(
						PlainAccess2 { canonicals[level + 1], canonical_level }, src, g);
			} else {
				k_downsample_average2d<<<counted(grid2(g)), block3(), 0, stream>>>(PackAccess2 { packs[level + 1], packs[level] },
						src, g);
				k_downsample_average2d<<<counted(grid2(g)), block3(), 0, stream>>>(
						PlainAccess2 { canonicals[level + 1], canonical_level }, src, g);
			}
			canonicals[level] = canonical_level;
		}
	}
	LSF_CUDA(cudaGetLastError());
	trace_point("pyramid");

	float *warp_current, *warp_next, *g_post, *scratch_a;
	unsigned* max_sq_bits;
	LSF_TRY(arena.alloc(&warp_current, N * 2));
	LSF_TRY(arena.alloc(&warp_next, N * 2));
	LSF_TRY(arena.alloc(&g_post, N * 2));
	LSF_TRY(arena.alloc(&scratch_a, N * 2));
	// one row of maximum slots per level: levels that run as ONE launch (termination test inside) are enqueued back to
	// back and the host reads all their slots at the end instead of waiting for every level
	const int slot_count = std::max(plan.max_iterations, 1);
	LSF_TRY(arena.alloc(&max_sq_bits, (size_t) slot_count * L));
	LSF_CUDA(cudaMemsetAsync(max_sq_bits, 0, (size_t) slot_count * L * sizeof(unsigned), stream));
	std::vector<unsigned> host_bits((size_t) slot_count * L);
	std::vector<char> deferred((size_t) L, 0);
	LSF_CUDA(cudaMemsetAsync(warp_current, 0, (size_t) plan.level_grid[0].N * 2 * sizeof(float), stream));
	TelemetryScratch telemetry;
	if (sink != nullptr) LSF_TRY(telemetry.allocate(arena, N, 2, plan.tikhonov, sink->want_fields != 0));

	float* capture_dev = nullptr;
	if (capture) {
		capture->count = 0;
		if (capture->level >= 0 && capture->level < L && capture->max_iterations > 0 && capture->buffer) {
			if (memory_kind == LSF_HOST)
				LSF_TRY(arena.alloc(&capture_dev, (size_t) capture->max_iterations * plan.level_grid[capture->level].N * 2));
			else capture_dev = capture->buffer;
		}
	}

	for (int level = 0; level < L; level++) {
		LevelState2 s;
		s.g = plan.level_grid[level];
		s.pack = packs[level];
		s.canonical = canonicals[level];
		s.warp = warp_current;
		s.g_post = g_post;
		s.scratch_a = scratch_a;
		s.max_sq_bits = max_sq_bits + (size_t) level * slot_count;
		unsigned* level_bits = host_bits.data() + (size_t) level * slot_count;
		LSF_CUDA(cudaMemsetAsync(s.g_post, 0, (size_t) s.g.N * 2 * sizeof(float), stream));
		const bool capturing = capture_dev != nullptr && capture->level == level;
		int executed = 0, enqueued = 0;
		bool converged = false;
		float last_max = FLT_MAX;
		if (sink != nullptr) {
			LSF_TRY(run_level_with_telemetry(plan, s, level, sink, telemetry, stream, &executed, &last_max));
			converged = true;  // the level is done
		}
		// small fields: the whole level in one cooperative launch (hier2d_persistent.cu); LSF_HIER2D_PERSISTENT=0 keeps one
		// launch per kernel (A/B)
		const char* persistent_env = getenv("LSF_HIER2D_PERSISTENT");
		const bool persistent = !capturing && sink == nullptr && !(persistent_env && persistent_env[0] == '0')
				&& s.g.N <= hier2d_persistent_capacity();
		// nothing the host does before the next level depends on this level's iteration count: look at the end
		const char* defer_env = getenv("LSF_HIER2D_DEFER_POLL");
		deferred[level] = persistent && !collect_reports && plan.max_iterations > 0 && !(defer_env && defer_env[0] == '0');
		if (deferred[level]) {
			LSF_TRY(enqueue_level_persistent(plan, s, 0, plan.max_iterations, stream));
			trace_point("enqueued");
			converged = true;  // (skips the polling loop)
		}
		while (!converged && enqueued < plan.max_iterations) {
			// the launch ends by itself at the first converged iteration: the host looks once per level
			const int chunk_end = persistent ? plan.max_iterations : std::min(plan.max_iterations, enqueued + POLL_CHUNK);
			if (persistent) LSF_TRY(enqueue_level_persistent(plan, s, enqueued, chunk_end - enqueued, stream));
			for (int it = enqueued; it < chunk_end && !persistent; it++) {
				enqueue_iteration(plan, s, it, stream);
				if (capturing && it < capture->max_iterations)
					k_planes_to_aos<<<counted(div_up(s.g.N, 256)), 256, 0, stream>>>(s.warp, capture_dev + (size_t) it * s.g.N * 2,
							s.g.N, 2);
			}
			LSF_CUDA(cudaGetLastError());
			LSF_CUDA(cudaMemcpyAsync(level_bits + enqueued, s.max_sq_bits + enqueued,
					(size_t) (chunk_end - enqueued) * sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
			trace_point("enqueued");
			LSF_CUDA(cudaStreamSynchronize(stream));
			for (int it = enqueued; it < chunk_end; it++) {
				float sq;
				std::memcpy(&sq, &level_bits[it], sizeof(float));
				if (it == enqueued) trace_point("level waited");
				last_max = std::sqrt(sq);
				executed = it + 1;
				if (last_max < plan.threshold) {
					converged = true;
					break;
				}
			}
			enqueued = chunk_end;
		}
		g_post = s.g_post;  // ping-pong state carries over (contents are reset per level)
		scratch_a = s.scratch_a;
		if (reports) {
			lsf_level_report& r = reports[level];
			std::memset(&r, 0, sizeof(r));
			r.iteration_count = executed;
			r.iteration_limit_reached = executed >= plan.max_iterations;
			r.max_update_length = last_max;
			r.dims[0] = s.g.H;
			r.dims[1] = s.g.W;
			r.dims[2] = 1;
			if (collect_reports) {
				// reference optimizer_with_telemetry.tpp:107-124
				float* live_level;
				LSF_TRY(arena.alloc(&live_level, (size_t) s.g.N));
				k_unpack_live2d<<<counted(grid2(s.g)), block3(), 0, stream>>>(s.pack, live_level, s.g);
				lsf_warp_delta_statistics_t w;
				lsf_tsdf_difference_statistics_t d;
				LSF_TRY(statistics_on_device(2, r.dims, s.warp, s.g.N, 1, s.canonical, live_level, plan.threshold, FLT_MAX, &w,
						&d, arena, stream));
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
		}
		if (capturing) capture->count = std::min(executed, capture->max_iterations);
		if (level != L - 1) {
			const Grid2& dg = plan.level_grid[level + 1];
			k_upsample2d<<<counted(grid2(dg)), block3(), 0, stream>>>(warp_current, warp_next, 2, s.g, dg, plan.linear ? 1 : 0);
			std::swap(warp_current, warp_next);
		}
	}
	k_planes_to_aos<<<counted(div_up(finest.N, 256)), 256, 0, stream>>>(warp_current, out_dev, finest.N, 2);
	LSF_CUDA(cudaGetLastError());
	trace_point("levels done");
	// the levels whose slots the host has not looked at yet: iteration counts and last maxima for the reports
	bool any_deferred = false;
	for (int level = 0; level < L; level++) any_deferred = any_deferred || deferred[level];
	if (any_deferred) {
		LSF_CUDA(cudaMemcpyAsync(host_bits.data(), max_sq_bits, (size_t) slot_count * L * sizeof(unsigned), cudaMemcpyDeviceToHost,
				stream));
		LSF_CUDA(cudaStreamSynchronize(stream));
		for (int level = 0; level < L && reports != nullptr; level++) {
			if (!deferred[level]) continue;
			int executed = 0;
			float last_max = FLT_MAX;
			for (int it = 0; it < plan.max_iterations; it++) {
				float sq;
				std::memcpy(&sq, &host_bits[(size_t) level * slot_count + it], sizeof(float));
				last_max = std::sqrt(sq);
				executed = it + 1;
				if (last_max < plan.threshold) break;  // reference optimizer.tpp:149,166-171
			}
			reports[level].iteration_count = executed;
			reports[level].iteration_limit_reached = executed >= plan.max_iterations;
			reports[level].max_update_length = last_max;
		}
		trace_point("slots read");
	}
	if (memory_kind == LSF_HOST) {
		if (capture_dev)
			LSF_CUDA(cudaMemcpyAsync(capture->buffer, capture_dev,
					(size_t) capture->count * plan.level_grid[capture->level].N * 2 * sizeof(float),
					cudaMemcpyDeviceToHost, stream));
		LSF_TRY(from_device(out_dev, warp_out, N * 2, LSF_HOST, stream));
	}
	trace_point("result");
	trace_point("end");
	return L;
}

}  // namespace

}  // namespace lsf

using namespace lsf;

extern "C" int lsf_hier_optimize_2d(const lsf_hier_params* params, const float* canonical, const float* live, int H,
		int W, float* warp_out, int memory_kind, lsf_level_report* reports, int collect_reports,
		lsf_iteration_capture* capture, void* stream_handle) {
	return hier_optimize_2d(params, canonical, live, H, W, warp_out, memory_kind, reports, collect_reports, capture, nullptr,
			stream_handle);
}

extern "C" int lsf_hier_optimize_2d_telemetry(const lsf_hier_params* params, const float* canonical, const float* live,
		int H, int W, float* warp_out, int memory_kind, lsf_level_report* reports, int collect_reports,
		const lsf_iteration_sink* sink, void* stream_handle) {
	LSF_REQUIRE(sink == nullptr || sink->callback != nullptr, "the iteration sink has no callback");
	return hier_optimize_2d(params, canonical, live, H, W, warp_out, memory_kind, reports, collect_reports, nullptr, sink,
			stream_handle);
}
