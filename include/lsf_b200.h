/*
 * lsf_b200.h -- C-ABI of liblsf_b200.so, the B200 (sm_100a) implementation of LevelSetFusion's
 * non-rigid warp-field optimisation hot path.
 *
 * This is the drop-in boundary: the entry points below are what the reference's Boost.Python module
 * `level_set_fusion_optimization` (reference cpp/src/module.cpp:26-50) binds for this path. Plain
 * pointers and sizes only; no torch / numpy / Eigen types. INTEGRATION.md shows the ctypes stub.
 *
 * Array conventions (identical to what the reference's numpy converters accept,
 * reference cpp/src/python_export/eigen_numpy_tensor.cpp:120-189, eigen_numpy_matrix.cpp:197-275):
 *   float32, C-contiguous, numpy index order.
 *   2D scalar f[H][W]; 2D vector v[H][W][2], component 0 = u (along columns), 1 = v (along rows).
 *   3D scalar f[X][Y][Z]; 3D vector v[X][Y][Z][3], component c displaces along axis c.
 *
 * Memory kinds: every entry point takes `memory_kind`: LSF_HOST = the pointers are host memory (the
 * library stages them through the device: this is the call the reference-facing Python shim makes
 * for numpy arguments), LSF_DEVICE = the pointers are device memory on the current device.
 * `stream` is a cudaStream_t passed as void* (NULL = default stream). All functions return 0 on
 * success or a negative lsf_status; lsf_last_error() describes the last failure of the calling thread.
 *
 * Error behaviour mirrors the reference's throw_assert preconditions
 * (reference cpp/src/nonrigid_optimization/hierarchical/pyramid.tpp:53-60, cpp/src/math/resampling.tpp:361,424,547).
 */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif

#define LSF_HOST 0
#define LSF_DEVICE 1
#define LSF_MAX_KERNEL_SIZE 31
#define LSF_MAX_LEVELS 16

typedef enum {
	LSF_OK = 0,
	LSF_ERR_INVALID_ARGUMENT = -1,   /* reference: AssertionFailureException -> RuntimeError */
	LSF_ERR_CUDA = -2,
	LSF_ERR_UNSUPPORTED = -3
} lsf_status;

/* reference: Optimizer<..>::ResamplingStrategy, cpp/src/nonrigid_optimization/hierarchical/optimizer.hpp:47-50 */
#define LSF_RESAMPLING_NEAREST_AND_AVERAGE 0
#define LSF_RESAMPLING_LINEAR 1

/* reference: constructor arguments of HierarchicalOptimizer2d/3d,
 * cpp/src/python_export/hierarchical_optimizer.tpp:47-71, defaults optimizer.hpp:51-65 */
typedef struct {
	int tikhonov_term_enabled;
	int gradient_kernel_enabled;
	int maximum_chunk_size;
	float rate;
	int maximum_iteration_count;
	float maximum_warp_update_threshold;
	float data_term_amplifier;
	float tikhonov_strength;
	const float* kernel; /* host pointer, kernel_size taps (odd, <= LSF_MAX_KERNEL_SIZE), may be NULL */
	int kernel_size;
	int resampling_strategy;
} lsf_hier_params;

/* per-level outcome; the statistics mirror reference cpp/src/telemetry/convergence_report.hpp:40-44,
 * warp_delta_statistics.hpp, tsdf_difference_statistics.hpp (filled when collect_reports != 0) */
typedef struct {
	int iteration_count;
	int iteration_limit_reached;
	float max_update_length;        /* last max ||gradient|| of the level */
	int dims[3];
	/* WarpDeltaStatistics over the band union */
	float warp_ratio_above_min_threshold, warp_length_min, warp_length_max, warp_length_mean, warp_length_std;
	int warp_longest_location[3];
	int warp_is_largest_below_min_threshold, warp_is_largest_above_max_threshold;
	/* TsdfDifferenceStatistics */
	float diff_min, diff_max, diff_mean, diff_std;
	int diff_biggest_location[3];
} lsf_level_report;

/* optional capture of the warp field after every iteration of one level (parity tests; the reference
 * offers the same through LoggingParameters.collect_per_level_iteration_data) */
typedef struct {
	int level;           /* 0 = coarsest; -1 = off */
	int max_iterations;
	float* buffer;       /* same memory kind as the call; [max_iterations][level voxels][D] */
	int count;           /* out */
} lsf_iteration_capture;

const char* lsf_last_error(void);
int lsf_version(void);
/* number of CUDA kernels this library has launched since it was loaded (bench.py's gpu_launches) */
long long lsf_launch_count(void);
/* Scratch memory comes from a memory pool owned by the library (one per device; freed blocks stay cached for the next
 * call; LSF_POOL_KEEP_MB bounds the cache). lsf_trim() synchronises the current device and returns every cached
 * block to the driver. The reference has no counterpart: its Eigen temporaries are heap allocations per call
 * (cpp/src/nonrigid_optimization/hierarchical/optimizer.tpp:186-207). */
int lsf_trim(void);

/* Debug / test introspection: kernel family used by the last 3D hierarchical iteration enqueued by this process
 * (the parity tests assert that they ran the kernels bench.py measures). */
#define LSF_PATH_OTHER 1            /* earlier kernel generations (A/B switches, unsupported shapes) */
#define LSF_PATH_TMA_STAGE1 2       /* k_hier_stage1_tma */
#define LSF_PATH_DEFERRED_UPDATE 4  /* ... with the warp update deferred into the next stage 1 (APPLY) */
#define LSF_PATH_FUSED_UPDATE 8     /* ... as the whole iteration (no Sobolev kernel configured) */
#define LSF_PATH_YMARCH2 16         /* k_sobolev_ymarch2 */
#define LSF_PATH_YMARCH3 32         /* k_sobolev_ymarch3 */
int lsf_debug_last_path(void);

/* ---------------------------------------------------------------- hierarchical optimizer
 * reference: Optimizer<S,V>::optimize(canonical_field, live_field) -> warp field,
 * cpp/src/nonrigid_optimization/hierarchical/optimizer.tpp:83-131 (level loop :134-171, iteration :174-212).
 * reports: array of LSF_MAX_LEVELS entries or NULL. Returns the level count (>0) on success. */
int lsf_hier_optimize_3d(const lsf_hier_params* params, const float* canonical, const float* live,
		int X, int Y, int Z, float* warp_out, int memory_kind, lsf_level_report* reports, int collect_reports,
		lsf_iteration_capture* capture, void* stream);
int lsf_hier_optimize_2d(const lsf_hier_params* params, const float* canonical, const float* live,
		int H, int W, float* warp_out, int memory_kind, lsf_level_report* reports, int collect_reports,
		lsf_iteration_capture* capture, void* stream);

/* Per-iteration telemetry (reference OptimizerWithTelemetry::optimize_iteration / optimize_level,
 * cpp/src/nonrigid_optimization/hierarchical/optimizer_with_telemetry.tpp:83-182): with a sink the optimizer runs one
 * iteration at a time and calls `callback` after every iteration with
 *   - the numbers behind the reference's VerbosityParameters prints (max update length, mean / standard deviation of
 *     diff = warped live - canonical, normalised data energy 1e6 * mean(diff^2), normalised Tikhonov energy
 *     1e6 * 0.5 * mean((sum of the Jacobian entries of the previous gradient)^2)), computed by GPU reductions, and
 *   - when want_fields != 0, what the reference's OptimizationIterationData stores per iteration
 *     (cpp/src/telemetry/optimization_iteration_data.tpp; LoggingParameters.collect_per_level_iteration_data): the live
 *     pyramid level, the warp field after the iteration, the data-term gradient (before the amplifier) and the
 *     Tikhonov-term gradient (before the strength; NULL when the term is off) as HOST arrays in numpy layout
 *     ([dims] / [dims][D]) that are valid during the callback only.
 * iteration == -1 marks the reference's initial frame of hierarchy level 0 (tpp:90-99: zero fields).
 * The result of optimize() is the same as without a sink (bit for bit); only the speed is not. */
typedef struct {
	int level;                       /* 0 = coarsest */
	int iteration;                   /* 0-based; -1 = initial frame of level 0 */
	int dims[3];                     /* level dimensions (2D: H, W, 1) */
	float max_update_length;
	float mean_tsdf_difference, std_tsdf_difference;
	float normalized_data_energy, normalized_tikhonov_energy;
	const float* live_field;
	const float* warp_field;
	const float* data_term_gradient;
	const float* tikhonov_term_gradient;
} lsf_iteration_record;
typedef void (*lsf_iteration_callback)(void* user, const lsf_iteration_record* record);
typedef struct {
	lsf_iteration_callback callback;
	void* user;
	int want_fields;
	int want_statistics;
} lsf_iteration_sink;
int lsf_hier_optimize_3d_telemetry(const lsf_hier_params* params, const float* canonical, const float* live,
		int X, int Y, int Z, float* warp_out, int memory_kind, lsf_level_report* reports, int collect_reports,
		const lsf_iteration_sink* sink, void* stream);
int lsf_hier_optimize_2d_telemetry(const lsf_hier_params* params, const float* canonical, const float* live,
		int H, int W, float* warp_out, int memory_kind, lsf_level_report* reports, int collect_reports,
		const lsf_iteration_sink* sink, void* stream);

/* Batched form for independent frame pairs (reference loop run_hierarchical_optimizer3d_multipair.py:403-406):
 * pair p uses canonical + p*X*Y*Z etc. iteration_counts: [pair_count][LSF_MAX_LEVELS] or NULL. The pairs advance in
 * lockstep through the pyramid levels and every iteration kernel covers all pairs of a sub-batch (one launch sequence
 * per iteration for the whole batch: the coarse levels are launch-bound otherwise), with per-pair termination on the
 * device; results per pair are bit-identical to lsf_hier_optimize_3d. Settings the batched kernels do not cover
 * (LINEAR resampling, kernels wider than 7 taps, Z not a multiple of 4) fall back to a loop over the pairs. */
int lsf_hier_optimize_3d_batch(const lsf_hier_params* params, const float* canonical, const float* live,
		int pair_count, int X, int Y, int Z, float* warp_out, int memory_kind, int* iteration_counts, void* stream);

/* Fixed-iteration driver of the finest-level iteration kernels on device-resident data (bench.py's
 * roofline leg): runs `iterations` iterations without termination test and without pyramid.
 * elapsed_ms receives the CUDA-event time on `stream`; kernel_launches the launches issued; stage_ms (4 floats or
 * NULL) the accumulated CUDA-event time of each kernel of the iteration (gradient stage, filter passes 0,1,2). */
int lsf_hier_iterate_3d(const lsf_hier_params* params, const float* canonical_dev, const float* live_dev,
		int X, int Y, int Z, int iterations, float* elapsed_ms, int* kernel_launches, float* stage_ms,
		void* stream);

/* ---------------------------------------------------------------- slab decomposition of one large volume (device pointers only)
 * A rank owns planes [own_begin, own_end) of an allocation of `planes` planes along numpy axis 0 (owned planes plus
 * halo planes towards interior cuts; no halo at the volume border). The reference has no counterpart: it is the
 * partitioning of Optimizer<Tensor3f,Tensor3v3f>::optimize_iteration (optimizer.tpp:174-212) named by SURVEY.md 8(e);
 * arithmetic per voxel is unchanged, so the sharded result is bit-identical to the whole-volume one. The host driver
 * (levelsetfusion-python_b200/slab.py) exchanges the halo planes between the two phases of an iteration. */
typedef struct {
	int planes, Y, Z;          /* allocation geometry of canonical / warp / g fields */
	int own_begin, own_end;    /* owned plane range inside the allocation */
	int x_origin;              /* global plane index of allocation plane 0 */
	int X_global;              /* planes of the whole level */
	const void* pack;          /* float4 pack of global planes [pack_origin, pack_origin + pack_planes), padded by 2 */
	int pack_planes, pack_origin;
	int pack_interior_low, pack_interior_high; /* 1: that pack edge is a cut, not the volume border */
	const float* canonical;    /* [planes][Y][Z] */
	float* warp;               /* [3][planes][Y][Z], updated in place on the owned planes */
	float* g_post;             /* [3][planes][Y][Z] filtered gradient (previous iteration in, this iteration out) */
	float* g_pre;              /* [3][planes][Y][Z] gradient before the filter (phase 1 out, phase 2 in) */
	unsigned* max_sq_bits;     /* one slot per iteration: bits of the rank-local max ||g||^2 */
	int* violation;            /* set to 1 if a gather left the rank's pack region */
} lsf_slab_level;

/* phase 1: gather + data term + Tikhonov term on the owned planes -> g_pre (without a Sobolev kernel: the whole
 * iteration, result in g_pre, caller swaps g_pre / g_post); phase 2: separable filter + warp update + max norm on the
 * owned planes (reads the +-radius halo planes of g_pre). The level-termination test reads slot iteration-1, which
 * the driver must have max-reduced over the ranks. */
int lsf_hier_slab_iteration(const lsf_hier_params* params, const lsf_slab_level* level, int iteration, int phase,
		void* stream);
/* pack {live, gradient} of global planes [pack_origin, pack_origin + pack_planes) from the live planes
 * [live_origin, live_origin + live_planes) (must contain the pack planes +-1 inside the volume) */
int lsf_slab_pack_finest(const float* live_region, int live_planes, int live_origin, int X_global, int Y, int Z,
		void* pack, int pack_planes, int pack_origin, void* stream);
/* restrict x2 by averaging (reference downsampleX2_average, resampling.tpp:385-417) between two slab allocations whose
 * plane 0 sits at global planes src_origin / dst_origin of their levels; kind 0 = scalar field, 1 = pack (planes
 * counted without the padding). Writes dst planes [dst_begin, dst_end). */
int lsf_slab_restrict(int kind, const void* src, int src_planes, int src_origin, int src_Y, int src_Z, void* dst,
		int dst_planes, int dst_origin, int dst_begin, int dst_end, void* stream);
/* prolong x2 nearest (reference upsampleX2_nearest, resampling.tpp:103-126) of a 3-component plane field */
int lsf_slab_prolong_nearest(const float* src, int src_planes, int src_origin, int src_Y, int src_Z, float* dst,
		int dst_planes, int dst_origin, int dst_begin, int dst_end, void* stream);

/* ---------------------------------------------------------------- slab decomposition: halo exchange through peer memory
 * The ranks of one NVLink / NVSwitch box (one process per GPU) exchange the halo planes of the slab decomposition and the
 * 4-byte level-termination maximum by storing straight into each other's memory: one kernel per exchange on every rank
 * (csrc/slab_peer.cu), no host round trip, no library collective. Replaces the per-iteration NCCL send / recv pairs and the
 * all_reduce of the first slab driver; no reference counterpart.
 * Every rank makes ONE allocation with lsf_peer_alloc (cudaMalloc + cudaIpcGetMemHandle; zero-filled), publishes the
 * 64-byte handle to the other ranks (any channel: torch.distributed.all_gather_object in slab.py) and maps theirs with
 * lsf_peer_open. All allocations share one layout: the exchanged fields at fixed byte offsets, a mailbox of
 * LSF_SLAB_MAILBOX_BYTES at mailbox_offset. */
#define LSF_SLAB_MAX_PEERS 8
#define LSF_SLAB_MAILBOX_BYTES 512
#define LSF_PEER_HANDLE_BYTES 64
int lsf_peer_alloc(size_t bytes, void** pointer_out, unsigned char* handle_out /* [LSF_PEER_HANDLE_BYTES] or NULL */);
int lsf_peer_open(const unsigned char* handle, void** pointer_out);
int lsf_peer_close(void* pointer);  /* a pointer from lsf_peer_open */
int lsf_peer_free(void* pointer);   /* a pointer from lsf_peer_alloc */
typedef struct {
	int rank, world_size;               /* world_size <= LSF_SLAB_MAX_PEERS */
	void* base[LSF_SLAB_MAX_PEERS];     /* every rank's allocation as mapped in this process (base[rank] = own) */
	size_t mailbox_offset;              /* byte offset of the mailbox inside every allocation (16-byte aligned) */
} lsf_slab_peers;
/* One exchange, enqueued on `stream` of every rank in the same order with the same `sequence` (1, 2, 3, ... over the life of
 * the allocations): copies this rank's first / last `width` owned planes of the 3-component field at `field_offset`
 * (geometry: level->planes / Y / Z / own_begin / own_end) into the halo planes of the low / high neighbour's field at the
 * same offset (their allocation has low_planes / high_planes planes, the halo starts at plane low_destination_plane /
 * high_destination_plane), signals, and waits for the neighbours' planes. reduce_iteration >= 0: additionally replaces
 * level->max_sq_bits[reduce_iteration] by the maximum over all ranks (width may be 0: reduction only). The kernel ends when
 * everything this rank waits for has arrived; a wait of more than ~4 s sets the error flag (lsf_slab_exchange_error). */
int lsf_slab_exchange(const lsf_slab_peers* peers, const lsf_slab_level* level, size_t field_offset, int width,
		int low_planes, int low_destination_plane, int high_planes, int high_destination_plane, int reduce_iteration,
		unsigned sequence, void* stream);
/* Geometry of the two exchanges of an iteration: byte offsets of g_pre / g_post inside every rank's allocation (they must
 * be the level's g_pre / g_post pointers minus base[rank]) and the neighbours' allocations at this level. */
typedef struct {
	size_t pre_offset, post_offset;
	int low_planes, low_own_end;       /* low neighbour: planes of its allocation, first plane of its high halo */
	int high_planes, high_own_begin;   /* high neighbour: planes of its allocation, first owned plane (its low halo ends there) */
} lsf_slab_link;
/* Enqueues iterations [first_iteration, first_iteration + iteration_count) of a level on this rank: per iteration
 * phase 1, [with a Sobolev kernel: exchange of `radius` g_pre planes, phase 2,] exchange of one g_post plane (Tikhonov term
 * enabled) together with the reduction of the iteration's maximum. Without a Sobolev kernel g_pre / g_post swap roles after
 * every iteration (the caller applies the parity to its own pointers). Exchanges take the sequence numbers first_sequence,
 * first_sequence + 1, ...; *sequences_used receives their count. One call replaces 4 host calls per iteration. */
int lsf_hier_slab_iterations(const lsf_hier_params* params, const lsf_slab_level* level, const lsf_slab_peers* peers,
		const lsf_slab_link* link, int first_iteration, int iteration_count, unsigned first_sequence,
		unsigned* sequences_used, void* stream);
/* synchronises `stream` and reports whether a wait of this rank has timed out (1) */
int lsf_slab_exchange_error(const lsf_slab_peers* peers, int* error_out, void* stream);

/* ---------------------------------------------------------------- primitives (device or host pointers)
 * reference: warp / warp_with_replacement, cpp/src/nonrigid_optimization/field_warping.tpp:68-225 */
int lsf_warp_3d(const float* field, int channels, const float* warp, int X, int Y, int Z, float oob_value,
		float* out, int memory_kind, void* stream);
int lsf_warp_2d(const float* field, int channels, const float* warp, int H, int W, float oob_value,
		float* out, int memory_kind, void* stream);
/* reference: math::gradient, cpp/src/math/gradients.tpp:248-283 (2D), :438-495 (3D) */
int lsf_gradient_3d(const float* field, int X, int Y, int Z, float* out, int memory_kind, void* stream);
int lsf_gradient_2d(const float* field, int H, int W, float* out, int memory_kind, void* stream);
/* reference: math::laplacian, cpp/src/math/gradients.tpp:62-101 (2D), :106-172 (3D) */
int lsf_laplacian_3d(const float* vfield, int X, int Y, int Z, float* out, int memory_kind, void* stream);
int lsf_laplacian_2d(const float* vfield, int H, int W, float* out, int memory_kind, void* stream);
/* reference: math::convolve_with_kernel[_preserve_zeros], cpp/src/math/convolution.cpp:69-332 (in place) */
int lsf_convolve_3d(float* vfield, int X, int Y, int Z, const float* kernel, int kernel_size, int memory_kind,
		void* stream);
int lsf_convolve_2d(float* vfield, int H, int W, const float* kernel, int kernel_size, int preserve_zeros,
		int memory_kind, void* stream);
/* reference: math::downsampleX2 / upsampleX2, cpp/src/math/resampling.tpp:68-656; channels = 1 (scalar) or D */
int lsf_downsample_3d(const float* field, int channels, int X, int Y, int Z, int linear, float* out,
		int memory_kind, void* stream);
int lsf_upsample_3d(const float* field, int channels, int X, int Y, int Z, int linear, float* out,
		int memory_kind, void* stream);
int lsf_downsample_2d(const float* field, int channels, int H, int W, int linear, float* out, int memory_kind,
		void* stream);
int lsf_upsample_2d(const float* field, int channels, int H, int W, int linear, float* out, int memory_kind,
		void* stream);
/* reference: math::locate_max_norm, cpp/src/math/statistics.tpp:57-100 */
int lsf_max_norm(const float* vfield, int channels, long long count, float* max_norm_out, int memory_kind,
		void* stream);
/* the same with the location of the longest vector. vfield: [H][W][channels] (nd 2) or [X][Y][Z][channels] (nd 3);
 * coordinates_out[3]: nd 2 -> (x, y) as statistics.tpp:70-71 decodes them (column, row of a square field), nd 3 -> (x, y, z);
 * among equal maxima the element the reference's traversal order meets first */
int lsf_locate_max_norm(const float* vfield, int channels, int nd, const int* dims, float* max_norm_out,
		int* coordinates_out, int memory_kind, void* stream);

/* ---------------------------------------------------------------- SobolevFusion / KillingFusion ("slavcheva") optimizers
 * reference: SobolevOptimizer2d::optimize(live_field, canonical_field) -> warped live field,
 * cpp/src/nonrigid_optimization/slavcheva/sobolev_optimizer2d.cpp:71-138 (parameters: optimizer2d.hpp:59-78,
 * sobolev_optimizer2d.hpp:39-70), and the Python class SlavchevaOptimizer2d,
 * nonrigid_opt/slavcheva/slavcheva_optimizer2d.py:72-430.
 * semantics selects whose arithmetic / loop structure is reproduced:
 *   LSF_SEMANTICS_CPP            C++ SobolevOptimizer2d; with the Killing / level-set terms and in 3D it is the dimensional
 *                                generalisation documented in DESIGN.md (the reference has no such optimizer)
 *   LSF_SEMANTICS_PY_DIRECT      Python ComputeMethod.DIRECT (2D only)
 *   LSF_SEMANTICS_PY_VECTORIZED  Python ComputeMethod.VECTORIZED (2D only; Tikhonov, no level-set term) */
#define LSF_SEMANTICS_CPP 0
#define LSF_SEMANTICS_PY_DIRECT 1
#define LSF_SEMANTICS_PY_VECTORIZED 2
#define LSF_DATA_TERM_BASIC 0            /* reference dt.DataTermMethod.BASIC, data_term.py:44-47 */
#define LSF_DATA_TERM_THRESHOLDED_FDM 1
#define LSF_SMOOTHING_TIKHONOV 0         /* reference st.SmoothingTermMethod, smoothing_term.py:27-29 */
#define LSF_SMOOTHING_KILLING 1

typedef struct {
	int semantics;
	int data_term_method;
	int smoothing_term_method;
	int level_set_term_enabled;
	int sobolev_smoothing_enabled;
	float gradient_descent_rate;
	float data_term_weight;
	float smoothing_term_weight;
	float isomorphic_enforcement_factor;
	float level_set_term_weight;
	float maximum_warp_length_lower_threshold;
	float maximum_warp_length_upper_threshold;
	int maximum_iteration_count;
	int minimum_iteration_count;
	const float* sobolev_kernel; /* host pointer, odd size <= LSF_MAX_KERNEL_SIZE, may be NULL */
	int sobolev_kernel_size;
} lsf_slavcheva_params;

/* reference telemetry::WarpDeltaStatistics, cpp/src/telemetry/warp_delta_statistics.hpp; locations are (x, y[, z]) =
 * position along the axis of component 0, 1[, 2] */
typedef struct {
	float ratio_above_min_threshold, length_min, length_max, length_mean, length_standard_deviation;
	int longest_warp_location[3];
	int is_largest_below_min_threshold, is_largest_above_max_threshold;
} lsf_warp_delta_statistics_t;

/* reference telemetry::TsdfDifferenceStatistics, cpp/src/telemetry/tsdf_difference_statistics.hpp */
typedef struct {
	float difference_min, difference_max, difference_mean, difference_standard_deviation;
	int biggest_difference_location[3];
} lsf_tsdf_difference_statistics_t;

/* reference telemetry::ConvergenceReport, cpp/src/telemetry/convergence_report.hpp:40-44 */
typedef struct {
	int iteration_count;
	int iteration_limit_reached;
	float last_max_warp_length;
	int has_statistics; /* set when collect_statistics != 0 */
	lsf_warp_delta_statistics_t warp_delta_statistics;
	lsf_tsdf_difference_statistics_t tsdf_difference_statistics;
} lsf_slavcheva_report;

/* optimize(live, canonical): nd = 2 ([H][W], square) or 3 ([X][Y][Z]); dims in numpy order.
 * live_out [dims]; warp_out [dims][nd] or NULL (the warp field of the last iteration); max_warps: host array of
 * max_warps_capacity floats or NULL, receives the maximum warp length of every iteration; capture (level ignored)
 * optionally receives the warp field after each of the first max_iterations iterations. */
int lsf_slavcheva_optimize(const lsf_slavcheva_params* params, const float* live, const float* canonical, int nd,
		const int* dims, float* live_out, float* warp_out, int memory_kind, lsf_slavcheva_report* report,
		int collect_statistics, float* max_warps, int max_warps_capacity, lsf_iteration_capture* capture, void* stream);
/* the same with the reference's per-iteration log (SharedParameters.enable_warp_statistics_logging,
 * cpp/src/nonrigid_optimization/slavcheva/sobolev_optimizer2d.cpp:88-97): iteration_statistics[i] (host array of
 * iteration_statistics_capacity entries, or NULL) receives the warp statistics over the band union of (canonical,
 * warped live) after iteration i; the rows of get_warp_statistics_as_matrix() (:144-160).
 * iteration_energies (host array [iteration_energies_capacity][3] of doubles, or NULL) receives what the reference's Python
 * optimizer appends to its OptimizationLog every iteration (nonrigid_opt/slavcheva/slavcheva_optimizer2d.py:370-374:
 * data_energies, smoothing_energies, level_set_energies, each times its weight; GPU reductions over the band union of the
 * state the iteration starts from). 2D fields with the Python semantics; zeros otherwise (the C++ optimizer drops its
 * energies, sobolev_optimizer2d.cpp:121-138). */
int lsf_slavcheva_optimize_logged(const lsf_slavcheva_params* params, const float* live, const float* canonical, int nd,
		const int* dims, float* live_out, float* warp_out, int memory_kind, lsf_slavcheva_report* report,
		int collect_statistics, float* max_warps, int max_warps_capacity, lsf_iteration_capture* capture,
		lsf_warp_delta_statistics_t* iteration_statistics, int iteration_statistics_capacity, double* iteration_energies,
		int iteration_energies_capacity, void* stream);

/* reference warp_2d_advanced / warp_2d_advanced_warp_unchanged, cpp/src/nonrigid_optimization/field_warping.cpp:64-154
 * (exported as warp_field_advanced[_no_warp_change], python_export/slavcheva_optimizer.cpp:64-89) and its 3D form.
 * warp [dims][nd] is updated in place when modify_warp != 0 (zeroed where the new value snaps to +-1). */
int lsf_warp_advanced(const float* live, const float* canonical, float* warp, int nd, const int* dims,
		int band_union_only, int known_values_only, int substitute_original, float truncation_float_threshold,
		int modify_warp, float* live_out, int memory_kind, void* stream);

/* reference build_warp_delta_statistics_{2d,3d} / build_tsdf_difference_statistics_{2d,3d},
 * cpp/src/python_export/telemetry.tpp:50-145 */
int lsf_warp_delta_statistics(const float* warp, const float* canonical, const float* live, int nd, const int* dims,
		float min_threshold, float max_threshold, lsf_warp_delta_statistics_t* out, int memory_kind, void* stream);
int lsf_tsdf_difference_statistics(const float* canonical, const float* live, int nd, const int* dims,
		lsf_tsdf_difference_statistics_t* out, int memory_kind, void* stream);

/* ---------------------------------------------------------------- TSDF generation from a depth image
 * reference tsdf::FilteringMethod, cpp/src/tsdf/interpolation_method.hpp:39-46 (exported python_export/tsdf.cpp:44-50) */
#define LSF_TSDF_FILTER_NONE 0
#define LSF_TSDF_FILTER_BILINEAR_IMAGE_SPACE 1      /* reference: "Not yet implemented" -> LSF_ERR_INVALID_ARGUMENT */
#define LSF_TSDF_FILTER_BILINEAR_VOXEL_SPACE 2      /* reference: "Not yet implemented" -> LSF_ERR_INVALID_ARGUMENT */
#define LSF_TSDF_FILTER_EWA_IMAGE_SPACE 3
#define LSF_TSDF_FILTER_EWA_VOXEL_SPACE 4
#define LSF_TSDF_FILTER_EWA_VOXEL_SPACE_INCLUSIVE 5

/* reference tsdf::Parameters<Container>, cpp/src/tsdf/parameters.hpp:31-56 (exported as tsdf.Parameters2d / Parameters3d,
 * python_export/tsdf.cpp:52-82). 2D: array_offset / field_shape hold (x, y) = (image-x direction, depth direction). */
typedef struct {
	float depth_unit_ratio;        /* metres per depth unit, reference default 0.001 */
	float projection_matrix[9];    /* camera intrinsics, row-major */
	float near_clipping_distance;  /* metres, default 0.05 */
	int array_offset[3];           /* voxels, default -64 */
	int field_shape[3];            /* voxels, default 128 */
	float voxel_size;              /* metres, default 0.004 */
	int narrow_band_width_voxels;  /* default 20 */
	int filtering_method;          /* LSF_TSDF_FILTER_*; the reference calls the member interpolation_method */
	float smoothing_factor;        /* covariance scale of the EWA methods, default 1 */
} lsf_tsdf_params;

/* reference tsdf::Generator2d / Generator3d ::generate(depth_image, camera_pose, image_y_coordinate),
 * cpp/src/tsdf/generator_crtp.tpp:40-71 -> generator_matrix.tpp:33-238 (2D), generator_tensor.tpp:40-270 (3D).
 * depth_image: uint16 [rows][cols] (memory_kind says where it and field_out live); camera_pose: HOST float[16], 4x4
 * row-major; image_y_coordinate: the image row a 2D field is generated from (ignored for nd = 3).
 * field_out: nd = 3 float [shape.x][shape.y][shape.z]; nd = 2 float [shape.y][shape.x]. Voxels behind the near clipping
 * distance, outside the image or without a depth reading keep the default value 1. */
int lsf_tsdf_generate(const lsf_tsdf_params* params, const unsigned short* depth_image, int rows, int cols,
		const float* camera_pose, int image_y_coordinate, int nd, float* field_out, int memory_kind, void* stream);

/* ---------------------------------------------------------------- rigid SDF-2-SDF tracker, 2D
 * reference rigid_optimization::Sdf2SdfOptimizer2d(rate, maximum_iteration_count, tsdf_generation_parameters,
 * verbosity_parameters).optimize(image_y_coordinate, canonical_field, live_depth_image, eta, initial_camera_pose) -> 3 x 3
 * twist matrix, cpp/src/rigid_optimization/sdf_2_sdf_optimizer2d.cpp:24-124 (exported
 * python_export/sdf_2_sdf_optimizer.cpp:22-58).
 * canonical_field float [shape.y][shape.x] and live_depth_image uint16 [rows][cols] live where memory_kind says;
 * initial_camera_pose (host 4 x 4 or NULL) is accepted and ignored like in the reference. Host outputs:
 * twist_matrix_out float[9] row-major; twists_out / optimal_twists_out float[maximum_iteration_count][3] and
 * energies_out float[maximum_iteration_count] (the numbers behind the reference's verbosity prints) may be NULL. */
int lsf_sdf2sdf_optimize_2d(const lsf_tsdf_params* tsdf_generation_parameters, float rate, int maximum_iteration_count,
		int image_y_coordinate, const float* canonical_field, const unsigned short* live_depth_image, int rows, int cols,
		float eta, const float* initial_camera_pose, float* twist_matrix_out, float* twists_out,
		float* optimal_twists_out, float* energies_out, int memory_kind, void* stream);

#ifdef __cplusplus
}
#endif
