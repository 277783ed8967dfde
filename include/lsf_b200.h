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

/* Batched form for independent frame pairs (reference loop run_hierarchical_optimizer3d_multipair.py:403-406):
 * pair p uses canonical + p*X*Y*Z etc. iteration_counts: [pair_count][LSF_MAX_LEVELS] or NULL. */
int lsf_hier_optimize_3d_batch(const lsf_hier_params* params, const float* canonical, const float* live,
		int pair_count, int X, int Y, int Z, float* warp_out, int memory_kind, int* iteration_counts, void* stream);

/* Fixed-iteration driver of the finest-level iteration kernels on device-resident data (bench.py's
 * roofline leg): runs `iterations` iterations without termination test and without pyramid.
 * elapsed_ms receives the CUDA-event time on `stream`; kernel_launches the launches issued; stage_ms (4 floats or
 * NULL) the accumulated CUDA-event time of each kernel of the iteration (gradient stage, filter passes 0,1,2). */
int lsf_hier_iterate_3d(const lsf_hier_params* params, const float* canonical_dev, const float* live_dev,
		int X, int Y, int Z, int iterations, float* elapsed_ms, int* kernel_launches, float* stage_ms,
		void* stream);

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

#ifdef __cplusplus
}
#endif
