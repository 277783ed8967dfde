/*
 * lsf_oracle.h -- C-ABI of the CPU ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY. This is a from-scratch CPU restatement of the reference's
 * (Algomorph/LevelSetFusion-Python + cpp/ submodule) non-rigid warp-field optimisation path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it. The product (levelsetfusion-python_b200/) never links, imports or calls it.
 *
 * Conventions (all arrays float32, C-contiguous, numpy index semantics):
 *   2D scalar field  f[H][W]        (row = y = axis 0, col = x = axis 1)
 *   2D vector field  v[H][W][2]     component 0 = u/x (displaces along columns, axis 1),
 *                                   component 1 = v/y (displaces along rows,    axis 0)
 *   3D scalar field  f[X][Y][Z]     numpy [i][j][k] == Eigen tensor(i,j,k)
 *   3D vector field  v[X][Y][Z][3]  component c displaces along numpy axis c
 * (reference: cpp/src/nonrigid_optimization/field_warping.tpp:29-43,80-89,159-168,
 *  cpp/src/python_export/eigen_numpy_tensor.cpp:168-187)
 *
 * Build: g++ -O3 -fopenmp -ffp-contract=off (no FMA contraction: the reference CI build is plain
 * SSE2 float32, cpp/CMakeLists.txt:65-73).
 */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	int tikhonov_term_enabled;
	int gradient_kernel_enabled;
	int maximum_chunk_size;
	float rate;
	int maximum_iteration_count;
	float maximum_warp_update_threshold;
	float data_term_amplifier;
	float tikhonov_strength;
	const float* kernel; /* may be NULL */
	int kernel_size;
	int resampling_strategy; /* 0 = NEAREST_AND_AVERAGE, 1 = LINEAR */
} orc_hier_params;

/* Optional per-iteration dump of one level's warp field (after the update of every iteration). */
typedef struct {
	int level;           /* level to dump (0 = coarsest), -1 = none */
	int max_iterations;  /* capacity of buffer in iterations */
	float* buffer;       /* [max_iterations][level voxels][D] */
	int count;           /* out: iterations dumped */
} orc_iteration_dump;

/* ---- primitives, 2D ---- */
void orc_warp2d(const float* field, const float* warp, int H, int W, float* out);
void orc_warp2d_replacement(const float* field, int C, const float* warp, int H, int W, float replacement, float* out);
void orc_gradient2d(const float* field, int H, int W, float* out);
void orc_laplacian2d(const float* vfield, int C, int H, int W, float* out);
void orc_convolve2d(float* vfield, int C, int H, int W, const float* kernel, int K, int preserve_zeros);
int orc_downsample2d(const float* field, int C, int H, int W, int linear, float* out);
int orc_upsample2d(const float* field, int C, int H, int W, int linear, float* out);
float orc_max_norm2d(const float* vfield, int C, long n);

/* ---- primitives, 3D ---- */
void orc_warp3d(const float* field, const float* warp, int X, int Y, int Z, float* out);
void orc_warp3d_replacement(const float* field, int C, const float* warp, int X, int Y, int Z, float replacement,
		float* out);
void orc_gradient3d(const float* field, int X, int Y, int Z, float* out);
void orc_laplacian3d(const float* vfield, int C, int X, int Y, int Z, float* out);
void orc_convolve3d(float* vfield, int C, int X, int Y, int Z, const float* kernel, int K);
int orc_downsample3d(const float* field, int C, int X, int Y, int Z, int linear, float* out);
int orc_upsample3d(const float* field, int C, int X, int Y, int Z, int linear, float* out);

/* ---- hierarchical optimizer ---- */
/* returns number of levels (>0) or a negative error code; iteration_counts[level] receives the
 * iterations executed per level, max_update_lengths[level] the last max ||gradient||. */
int orc_hier_optimize2d(const orc_hier_params* p, const float* canonical, const float* live, int H, int W,
		float* warp_out, int* iteration_counts, float* max_update_lengths, orc_iteration_dump* dump);
int orc_hier_optimize3d(const orc_hier_params* p, const float* canonical, const float* live, int X, int Y, int Z,
		float* warp_out, int* iteration_counts, float* max_update_lengths, orc_iteration_dump* dump);

/* timing helper for bench.py's cpu_baseline: runs `iterations` fixed iterations of the finest-level
 * hierarchical step on a 3D pair (no pyramid), returns seconds of wall time (omp_get_wtime). */
double orc_hier_time_iterations3d(const orc_hier_params* p, const float* canonical, const float* live, int X, int Y,
		int Z, int iterations);

int orc_num_threads(void);
/* launchers such as torch.distributed.run export OMP_NUM_THREADS=1: lets the timing legs use every host core */
void orc_set_num_threads(int n);

/* ---- SobolevFusion / KillingFusion ("slavcheva") optimizers: lsf_oracle_slavcheva.cpp ---- */
#define ORC_SEMANTICS_CPP 0            /* C++ SobolevOptimizer2d (+ dimensional generalisation) */
#define ORC_SEMANTICS_PY_DIRECT 1      /* Python SlavchevaOptimizer2d, ComputeMethod.DIRECT */
#define ORC_SEMANTICS_PY_VECTORIZED 2  /* Python SlavchevaOptimizer2d, ComputeMethod.VECTORIZED */
#define ORC_DATA_TERM_BASIC 0
#define ORC_DATA_TERM_THRESHOLDED_FDM 1
#define ORC_SMOOTHING_TIKHONOV 0
#define ORC_SMOOTHING_KILLING 1

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
	int max_iterations;
	int min_iterations;
	const float* kernel; /* may be NULL */
	int kernel_size;
} orc_slavcheva_params;

typedef struct {
	float ratio_above_min_threshold, length_min, length_max, length_mean, length_standard_deviation;
	int longest_warp_location[3]; /* (x, y[, z]) = position along the axis of component 0, 1[, 2] */
	int is_largest_below_min_threshold, is_largest_above_max_threshold;
} orc_warp_delta_statistics_t;

typedef struct {
	float difference_min, difference_max, difference_mean, difference_standard_deviation;
	int biggest_difference_location[3];
} orc_tsdf_difference_statistics_t;

/* optimize(live, canonical) -> warped live field (+ the last warp field). dims: nd extents in numpy order.
 * max_warps[i] receives the maximum warp length reported by iteration i. dump (optional) receives the warp
 * field after every iteration (level member ignored). Returns 0 or a negative error code. */
int orc_slavcheva_optimize(const orc_slavcheva_params* p, const float* live, const float* canonical, int nd,
		const int* dims, float* live_out, float* warp_out, int* iteration_count, float* max_warps, int max_warps_capacity,
		orc_iteration_dump* dump);
/* the same; energies [energies_capacity][3] (or NULL) receives the {data, smoothing, level set} energy aggregates the
 * reference's Python optimizer logs for every iteration (2D, Python semantics; zeros otherwise) */
int orc_slavcheva_optimize_energies(const orc_slavcheva_params* p, const float* live, const float* canonical, int nd,
		const int* dims, float* live_out, float* warp_out, int* iteration_count, float* max_warps, int max_warps_capacity,
		orc_iteration_dump* dump, double* energies, int energies_capacity);
/* {data, smoothing, level set} energy aggregates of one state (live = warped live field, warp [dims][nd]) */
void orc_slavcheva_energies(const orc_slavcheva_params* p, const float* live, const float* canonical, const float* warp,
		int nd, const int* dims, double* out);
void orc_slavcheva_data_term(const orc_slavcheva_params* p, const float* live, const float* canonical, int nd,
		const int* dims, int band_union_only, float* out);
void orc_slavcheva_smoothing_term(const orc_slavcheva_params* p, const float* warp, const float* live,
		const float* canonical, int nd, const int* dims, int band_union_only, float* out);
void orc_slavcheva_level_set_term(const orc_slavcheva_params* p, const float* live, int nd, const int* dims, float* out);
/* warp_2d_advanced and its 3D generalisation; warp is updated in place when modify_warp != 0 */
void orc_warp_advanced(const float* live, const float* canonical, float* warp, int nd, const int* dims,
		int band_union_only, int known_values_only, int substitute_original, float truncation_float_threshold,
		int modify_warp, float* new_live);
void orc_warp_delta_statistics(const float* warp, const float* canonical, const float* live, int nd, const int* dims,
		float min_threshold, float max_threshold, orc_warp_delta_statistics_t* out);
void orc_tsdf_difference_statistics(const float* canonical, const float* live, int nd, const int* dims,
		orc_tsdf_difference_statistics_t* out);

/* ---------------------------------------------------------------- TSDF generation from a depth image (lsf_oracle_tsdf.cpp)
 * reference tsdf::Parameters, cpp/src/tsdf/parameters.hpp:31-56 (projection_matrix row-major) */
typedef struct {
	float depth_unit_ratio;
	float projection_matrix[9];
	float near_clipping_distance;
	int array_offset[3];
	int field_shape[3];
	float voxel_size;
	int narrow_band_width_voxels;
	int filtering_method; /* 0 = NONE (the only one restated) */
	float smoothing_factor;
} orc_tsdf_params;
/* nd 3: field [shape.x][shape.y][shape.z]; nd 2: field [shape.y][shape.x] from image row image_y_coordinate.
 * depth_image [rows][cols] uint16, pose 4x4 row-major. Returns 0, or -1 for an unsupported filtering method. */
int orc_tsdf_generate(const orc_tsdf_params* p, const unsigned short* depth_image, int rows, int cols, const float* pose,
		int image_y_coordinate, int nd, float* field);

/* ---------------------------------------------------------------- rigid SDF-2-SDF tracker, 2D (lsf_oracle_rigid.cpp)
 * reference Sdf2SdfOptimizer2d::optimize, cpp/src/rigid_optimization/sdf_2_sdf_optimizer2d.cpp:63-124.
 * canonical_field [shape.y][shape.x]; twist_matrix_out 3x3 row-major; twists_out [maximum_iteration_count][3] and
 * energies_out [maximum_iteration_count] may be NULL. Returns 0, or the TSDF generator's error. */
int orc_sdf2sdf_optimize(const orc_tsdf_params* tsdf_parameters, float rate, int maximum_iteration_count,
		int image_y_coordinate, const float* canonical_field, const unsigned short* live_depth_image, int rows, int cols,
		float eta, int double_sums, float* twist_matrix_out, float* twists_out, float* energies_out);

#ifdef __cplusplus
}
#endif
