// tsdf.cu -- projective TSDF generation from a depth image on the GPU (the step immediately
// before the optimisation path; it produces the canonical / live fields the optimizers consume).
//
// Replaces, with the reference's float32 operation order (compiled --fmad=false like the rest of the library):
//   tsdf::Generator3d::generate   reference cpp/src/tsdf/generator_tensor.tpp:40-101 (NONE), :126-270 (EWA)
//   tsdf::Generator2d::generate   reference cpp/src/tsdf/generator_matrix.tpp:33-93 (NONE), :118-238 (EWA)
//   helpers                       cpp/src/tsdf/common.hpp:29-51, ewa_common.hpp:32-235, cpp/src/math/conics.cpp:44-59
//
// Layout: the 3D field is [x][y][z] with z fastest (what the reference's column-major Eigen tensor becomes in numpy),
// the 2D field [y][x]. One thread owns VOXELS consecutive voxels along the fastest axis and writes them with one
// 128-bit store; the per-(x, y) part of the camera transform is computed once per thread. Voxels that the reference
// skips receive the default value 1 in the same store: the field is written exactly once (4 B per voxel; the depth
// image, 0.6 MB, stays in L2).
// Filtering NONE is bit-identical to the CPU oracle, and so are the EWA methods: their weights are evaluated with the
// host libm's expf algorithm (expf_as_host below).
#include "common.cuh"

#include <cfloat>
#include <cstring>

namespace lsf {
namespace {

struct TsdfArgs {
	float pose[12];        // rows 0..2 of the camera pose (row-major)
	float projection[6];   // rows 0..1 of the projection matrix
	float covariance[9];   // compute_covariance_camera_space
	float near_clipping_distance;
	float depth_unit_ratio;
	float voxel_size;
	float half_width;
	float threshold;       // squared_radius_threshold
	int offset[3];
	int shape[3];          // 2D: shape[0] = x (columns), shape[1] = y (rows of the field), shape[2] = 1
	int rows, cols;
	int image_y_coordinate;
	const unsigned short* __restrict__ depth;
	float* __restrict__ field;
};

// expf as the reference's std::exp(float) evaluates it on the host: glibc's expf (sysdeps/ieee754/flt-32/e_expf.c, the
// Arm optimized-routines algorithm, glibc >= 2.27): x * 32 / ln 2 = k + r, exp(x) = 2^(k / 32) * (C0 r^3 + C1 r^2 + C2 r + 1)
// in DOUBLE precision with a 32-entry table of 2^(i / 32), rounded to float once at the end. The EWA weights must be
// bit-identical to the host's: a one-ulp difference in a single weight moves the rounding of the float sums that follow,
// and the quotient of the sums is amplified by depth / narrow-band half-width (about 80) -- 2e-5 in TSDF units with the
// CUDA expf, against the 1e-6 the reference's own EWA tests allow. Checked against libm for every float in [-80, 0]
// (1.1e9 inputs, one mismatch) when this was written. Arguments outside (-80, 80) fall back to expf (never reached by
// the generators: the argument is -0.5 * dist_sq with dist_sq <= 4 * voxel_size * smoothing_factor).
__constant__ unsigned long long EXP2F_TABLE[32] = {
0x3ff0000000000000ULL,
0x3fefd9b0d3158574ULL,
0x3fefb5586cf9890fULL,
0x3fef9301d0125b51ULL,
0x3fef72b83c7d517bULL,
0x3fef54873168b9aaULL,
0x3fef387a6e756238ULL,
0x3fef1e9df51fdee1ULL,
0x3fef06fe0a31b715ULL,
0x3feef1a7373aa9cbULL,
0x3feedea64c123422ULL,
0x3feece086061892dULL,
0x3feebfdad5362a27ULL,
0x3feeb42b569d4f82ULL,
0x3feeab07dd485429ULL,
0x3feea47eb03a5585ULL,
0x3feea09e667f3bcdULL,
0x3fee9f75e8ec5f74ULL,
0x3feea11473eb0187ULL,
0x3feea589994cce13ULL,
0x3feeace5422aa0dbULL,
0x3feeb737b0cdc5e5ULL,
0x3feec49182a3f090ULL,
0x3feed503b23e255dULL,
0x3feee89f995ad3adULL,
0x3feeff76f2fb5e47ULL,
0x3fef199bdd85529cULL,
0x3fef3720dcef9069ULL,
0x3fef5818dcfba487ULL,
0x3fef7c97337b9b5fULL,
0x3fefa4afa2a490daULL,
0x3fefd0765b6e4540ULL};

__device__ __forceinline__ float expf_as_host(float x) {
	if (!(fabsf(x) < 80.0f)) return expf(x);
	const double N = 32.0;
	const double z = (0x1.71547652b82fep+0 * N) * (double) x;
	double kd = z + 0x1.8p+52;
	const unsigned long long ki = (unsigned long long) __double_as_longlong(kd);
	kd -= 0x1.8p+52;
	const double r = z - kd;
	const unsigned long long t = EXP2F_TABLE[ki & 31] + (ki << 47);
	const double s = __longlong_as_double((long long) t);
	const double q = (0x1.c6af84b912394p-5 / N / N / N) * r + (0x1.ebfce50fac4f3p-3 / N / N);
	const double r2 = r * r;
	double y = (0x1.62e42ff0c52d6p-1 / N) * r + 1.0;
	y = q * r2 + y;
	y = y * s;
	return (float) y;
}

__device__ __forceinline__ float tsdf_value(float signed_distance, float half_width) {
	// reference compute_TSDF_value, common.hpp:32-40
	return signed_distance < -half_width ? -1.0f : (signed_distance > half_width ? 1.0f : signed_distance / half_width);
}

// METHOD: LSF_TSDF_FILTER_*; ND: 2 or 3. The value of one voxel; 1 where the reference leaves the default.
template<int METHOD, int ND>
__device__ __forceinline__ float voxel_value(const TsdfArgs& a, float x_voxel, float y_voxel, float z_voxel) {
	float cam[3];
#pragma unroll
	for (int r = 0; r < 3; r++)
		cam[r] = ((a.pose[4 * r] * x_voxel + a.pose[4 * r + 1] * y_voxel) + a.pose[4 * r + 2] * z_voxel) + a.pose[4 * r + 3] * 1.0f;
	// the 2D EWA generators compare with the global constant instead of the parameter (generator_matrix.tpp:146)
	const float near = (ND == 2 && METHOD != LSF_TSDF_FILTER_NONE) ? 0.05f : a.near_clipping_distance;
	if (cam[2] <= near) return 1.0f;
	const float* P = a.projection;
	const float image_x = ((P[0] * cam[0] + P[1] * cam[1]) + P[2] * cam[2]) / cam[2];
	const float image_y = ND == 2 ? (float) a.image_y_coordinate : ((P[3] * cam[0] + P[4] * cam[1]) + P[5] * cam[2]) / cam[2];
	const int rows = a.rows, cols = a.cols;
	if (METHOD == LSF_TSDF_FILTER_NONE) {
		if (image_x < 0.0f || image_x >= (float) cols || image_y < 0.0f || image_y >= (float) rows) return 1.0f;
		// 3D rounds in double (generator_tensor.tpp:86-87), 2D in float (generator_matrix.tpp:77)
		const int ix = ND == 2 ? (int) (image_x + 0.5f) : (int) ((double) image_x + 0.5);
		const int iy = ND == 2 ? a.image_y_coordinate : (int) ((double) image_y + 0.5);
		if (ix >= cols || iy >= rows) return 1.0f;  // the reference reads outside the image here
		const float depth = (float) __ldg(a.depth + (size_t) iy * cols + ix) * a.depth_unit_ratio;
		if (depth <= 0.0f) return 1.0f;
		return tsdf_value(depth - cam[2], a.half_width);
	}
	// ---- elliptical weighted average (generate__ewa_aux)
	if (image_x < -3.0f || image_x >= (float) (cols + 3) || image_y < -3.0f || image_y >= (float) (rows + 3)) return 1.0f;
	const float ray_distance = sqrtf((cam[0] * cam[0] + cam[1] * cam[1]) + cam[2] * cam[2]);
	const float z_cam_squared = cam[2] * cam[2];
	const float inv_z_cam = 1.0f / cam[2];
	const float J[3][3] = { { inv_z_cam, 0.0f, -cam[0] / z_cam_squared }, { 0.0f, inv_z_cam, -cam[1] / z_cam_squared },
			{ cam[0] / ray_distance, cam[1] / ray_distance, cam[2] / ray_distance } };
	float T[2][3], R2[2][2];
#pragma unroll
	for (int i = 0; i < 2; i++)
#pragma unroll
		for (int j = 0; j < 3; j++)
			T[i][j] = (J[i][0] * a.covariance[j] + J[i][1] * a.covariance[3 + j]) + J[i][2] * a.covariance[6 + j];
#pragma unroll
	for (int i = 0; i < 2; i++)
#pragma unroll
		for (int j = 0; j < 2; j++) R2[i][j] = (T[i][0] * J[j][0] + T[i][1] * J[j][1]) + T[i][2] * J[j][2];
	const float S[2][2] = { { P[0], P[1] }, { P[3], P[4] } };
	float SR[2][2], F2[2][2];
#pragma unroll
	for (int i = 0; i < 2; i++)
#pragma unroll
		for (int j = 0; j < 2; j++) SR[i][j] = S[i][0] * R2[0][j] + S[i][1] * R2[1][j];
#pragma unroll
	for (int i = 0; i < 2; i++)
#pragma unroll
		for (int j = 0; j < 2; j++) F2[i][j] = (SR[i][0] * S[j][0] + SR[i][1] * S[j][1]) + (i == j ? 1.0f : 0.0f);
	const float determinant = F2[0][0] * F2[1][1] - F2[1][0] * F2[0][1];
	const float inverse_determinant = 1.0f / determinant;
	const float Q00 = F2[1][1] * inverse_determinant, Q01 = -F2[0][1] * inverse_determinant;
	const float Q10 = -F2[1][0] * inverse_determinant, Q11 = F2[0][0] * inverse_determinant;
	// compute_centered_ellipse_bound_points (conics.cpp:44-59; the tilted branch exchanges the axes, kept)
	const float B = Q01 * 2.0f, F = a.threshold;
	float bound_x, bound_y;
	if (fabsf(B) < FLT_EPSILON) {
		bound_x = sqrtf(F / Q00);
		bound_y = sqrtf(F / Q11);
	} else {
		const float B_squared = B * B;
		bound_x = sqrtf(F / (Q11 - B_squared / (4.0f * Q00)));
		bound_y = sqrtf(F / (Q00 - B_squared / (4.0f * Q11)));
	}
	int x_start = (int) (image_x - bound_x);
	int x_end = (int) ceilf(image_x + bound_x + 1.0f);
	int y_start = (int) (image_y - bound_y);
	int y_end = (int) ceilf(image_y + bound_y + 1.0f);
	if (x_start >= cols || x_end <= 0 || y_start >= rows || y_end <= 0) return 1.0f;
	if (METHOD != LSF_TSDF_FILTER_EWA_VOXEL_SPACE_INCLUSIVE) {
		x_start = max(x_start, 0);
		x_end = min(x_end, cols);
		y_start = max(y_start, 0);
		y_end = min(y_end, rows);
	}
	float weights_sum = 0.0f, value_sum = 0.0f;
	for (int x_sample = x_start; x_sample < x_end; x_sample++) {
		const float sx = (float) x_sample - image_x;
		for (int y_sample = y_start; y_sample < y_end; y_sample++) {
			const float sy = (float) y_sample - image_y;
			const float dist_sq = (sx * Q00 + sy * Q10) * sx + (sx * Q01 + sy * Q11) * sy;
			if (dist_sq > F) continue;
			const float weight = expf_as_host(-0.5f * dist_sq);
			if (METHOD == LSF_TSDF_FILTER_EWA_VOXEL_SPACE_INCLUSIVE
					&& (y_sample < 0 || y_sample >= rows || x_sample < 0 || x_sample >= cols)) {
				value_sum += weight;
				weights_sum += weight;
				continue;
			}
			const float surface_depth = (float) __ldg(a.depth + (size_t) y_sample * cols + x_sample) * a.depth_unit_ratio;
			if (surface_depth <= 0.0f) continue;
			if (METHOD == LSF_TSDF_FILTER_EWA_IMAGE_SPACE) value_sum += weight * surface_depth;
			else value_sum += weight * tsdf_value(surface_depth - cam[2], a.half_width);
			weights_sum += weight;
		}
	}
	if (METHOD == LSF_TSDF_FILTER_EWA_IMAGE_SPACE) {
		if (value_sum <= 0.0f) return 1.0f;
		return tsdf_value(value_sum / weights_sum - cam[2], a.half_width);
	}
	return weights_sum == 0.0f ? 1.0f : value_sum / weights_sum;
}

// 3D: one thread = VOXELS consecutive z of one (x, y) column; the grid covers (x * sy + y, z / VOXELS) with z fastest so
// that a warp writes 32 * VOXELS * 4 contiguous bytes. 2D: one thread = VOXELS consecutive x of one field row.
template<int METHOD, int ND, int VOXELS>
__global__ void __launch_bounds__(256) k_tsdf_generate(const TsdfArgs a) {
	const int fast_extent = ND == 3 ? a.shape[2] : a.shape[0];
	const int groups_per_line = (fast_extent + VOXELS - 1) / VOXELS;
	const long long lines = ND == 3 ? (long long) a.shape[0] * a.shape[1] : a.shape[1];
	const long long total = lines * groups_per_line;
	for (long long g = blockIdx.x * (long long) blockDim.x + threadIdx.x; g < total; g += (long long) gridDim.x * blockDim.x) {
		const long long line = g / groups_per_line;
		const int first = (int) (g - line * groups_per_line) * VOXELS;
		float values[VOXELS];
		if (ND == 3) {
			const int x_field = (int) (line / a.shape[1]), y_field = (int) (line - (long long) x_field * a.shape[1]);
			const float x_voxel = (float) (x_field + a.offset[0]) * a.voxel_size;
			const float y_voxel = (float) (y_field + a.offset[1]) * a.voxel_size;
#pragma unroll
			for (int v = 0; v < VOXELS; v++) {
				const float z_voxel = (float) (first + v + a.offset[2]) * a.voxel_size;
				values[v] = first + v < fast_extent ? voxel_value<METHOD, 3>(a, x_voxel, y_voxel, z_voxel) : 1.0f;
			}
		} else {
			const float z_voxel = (float) ((int) line + a.offset[1]) * a.voxel_size;
#pragma unroll
			for (int v = 0; v < VOXELS; v++) {
				const float x_voxel = (float) (first + v + a.offset[0]) * a.voxel_size;
				values[v] = first + v < fast_extent ? voxel_value<METHOD, 2>(a, x_voxel, 0.0f, z_voxel) : 1.0f;
			}
		}
		float* out = a.field + line * fast_extent + first;
		if (VOXELS == 4 && first + 4 <= fast_extent && (fast_extent & 3) == 0) {
			*reinterpret_cast<float4*>(out) = make_float4(values[0], values[1], values[2], values[3]);
		} else {
#pragma unroll
			for (int v = 0; v < VOXELS; v++)
				if (first + v < fast_extent) out[v] = values[v];
		}
	}
}

template<int ND>
cudaError_t launch(const TsdfArgs& a, int method, cudaStream_t stream) {
	const int fast_extent = ND == 3 ? a.shape[2] : a.shape[0];
	const long long lines = ND == 3 ? (long long) a.shape[0] * a.shape[1] : a.shape[1];
	int sm_count = 148;
	int device = 0;
	cudaGetDevice(&device);
	cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device);
	// NONE: four voxels per thread (one 128-bit store), grid = whole waves of 8 blocks per SM, grid-stride beyond that.
	// EWA: one voxel per thread -- the sample loops of neighbouring voxels have similar trip counts, and a thread's
	// time is the loop, not the store.
	if (method == LSF_TSDF_FILTER_NONE) {
		const long long groups = lines * ((fast_extent + 3) / 4);
		const unsigned blocks = (unsigned) std::min<long long>((groups + 255) / 256, (long long) sm_count * 8 * 4);
		k_tsdf_generate<LSF_TSDF_FILTER_NONE, ND, 4> <<<counted(std::max(blocks, 1u)), 256, 0, stream>>>(a);
		return cudaGetLastError();
	}
	const long long groups = lines * fast_extent;
	const unsigned blocks = (unsigned) std::max<long long>(std::min<long long>((groups + 255) / 256, (long long) sm_count * 8 * 16), 1);
	switch (method) {
	case LSF_TSDF_FILTER_EWA_IMAGE_SPACE:
		k_tsdf_generate<LSF_TSDF_FILTER_EWA_IMAGE_SPACE, ND, 1> <<<counted(blocks), 256, 0, stream>>>(a);
		break;
	case LSF_TSDF_FILTER_EWA_VOXEL_SPACE:
		k_tsdf_generate<LSF_TSDF_FILTER_EWA_VOXEL_SPACE, ND, 1> <<<counted(blocks), 256, 0, stream>>>(a);
		break;
	default:
		k_tsdf_generate<LSF_TSDF_FILTER_EWA_VOXEL_SPACE_INCLUSIVE, ND, 1> <<<counted(blocks), 256, 0, stream>>>(a);
		break;
	}
	return cudaGetLastError();
}

}  // namespace
}  // namespace lsf

namespace lsf {

// the generator on device buffers (depth image and field already on the device); shared with the rigid tracker
int tsdf_generate_device(const lsf_tsdf_params* params, const unsigned short* depth_dev, int rows, int cols,
		const float* camera_pose, int image_y_coordinate, int nd, float* field_dev, cudaStream_t stream) {
	LSF_REQUIRE(nd == 2 || nd == 3, "nd must be 2 or 3, got %d", nd);
	LSF_REQUIRE(rows > 0 && cols > 0, "depth image must not be empty, got %d x %d", rows, cols);
	const int method = params->filtering_method;
	// reference generator_matrix.tpp:95-115, generator_tensor.tpp:103-123: throw_assert(false, "Not yet implemented")
	LSF_REQUIRE(method != LSF_TSDF_FILTER_BILINEAR_IMAGE_SPACE && method != LSF_TSDF_FILTER_BILINEAR_VOXEL_SPACE,
			"Not yet implemented");
	// reference generator_crtp.tpp:66-70
	LSF_REQUIRE(method == LSF_TSDF_FILTER_NONE || method == LSF_TSDF_FILTER_EWA_IMAGE_SPACE
			|| method == LSF_TSDF_FILTER_EWA_VOXEL_SPACE || method == LSF_TSDF_FILTER_EWA_VOXEL_SPACE_INCLUSIVE,
			"Unknown InterpolationMethod enum value, %d", method);
	TsdfArgs a;
	std::memset(&a, 0, sizeof(a));
	for (int d = 0; d < 3; d++) {
		a.shape[d] = d < nd ? params->field_shape[d] : 1;
		a.offset[d] = d < nd ? params->array_offset[d] : 0;
		LSF_REQUIRE(a.shape[d] > 0, "field_shape[%d] must be positive, got %d", d, a.shape[d]);
	}
	if (nd == 2)
		LSF_REQUIRE(image_y_coordinate >= 0 && image_y_coordinate < rows, "image_y_coordinate %d outside the %d image rows",
				image_y_coordinate, rows);
	std::memcpy(a.pose, camera_pose, sizeof(a.pose));
	std::memcpy(a.projection, params->projection_matrix, sizeof(a.projection));
	a.near_clipping_distance = params->near_clipping_distance;
	a.depth_unit_ratio = params->depth_unit_ratio;
	a.voxel_size = params->voxel_size;
	a.half_width = (float) (((double) (float) params->narrow_band_width_voxels / 2.) * (double) params->voxel_size);
	a.threshold = 4.0f * params->voxel_size * params->smoothing_factor;
	{
		// compute_covariance_camera_space (ewa_common.hpp:32-42): R * (I * (voxel_size * scale)) * R^T
		const float s = params->voxel_size * params->smoothing_factor;
		float M[3][3];
		for (int i = 0; i < 3; i++)
			for (int j = 0; j < 3; j++) M[i][j] = camera_pose[4 * i + j] * s;
		for (int i = 0; i < 3; i++)
			for (int j = 0; j < 3; j++)
				a.covariance[3 * i + j] = (M[i][0] * camera_pose[4 * j] + M[i][1] * camera_pose[4 * j + 1])
						+ M[i][2] * camera_pose[4 * j + 2];
	}
	a.rows = rows;
	a.cols = cols;
	a.image_y_coordinate = image_y_coordinate;
	a.depth = depth_dev;
	a.field = field_dev;
	LSF_CUDA(nd == 3 ? launch<3>(a, method, stream) : launch<2>(a, method, stream));
	return LSF_OK;
}

}  // namespace lsf

using namespace lsf;

extern "C" int lsf_tsdf_generate(const lsf_tsdf_params* params, const unsigned short* depth_image, int rows, int cols,
		const float* camera_pose, int image_y_coordinate, int nd, float* field_out, int memory_kind, void* stream_handle) {
	cudaStream_t stream = static_cast<cudaStream_t>(stream_handle);
	LSF_REQUIRE(params && depth_image && camera_pose && field_out, "params, depth_image, camera_pose and field_out must not be NULL");
	LSF_REQUIRE(nd == 2 || nd == 3, "nd must be 2 or 3, got %d", nd);
	LSF_REQUIRE(rows > 0 && cols > 0, "depth image must not be empty, got %d x %d", rows, cols);
	size_t N = 1;
	for (int d = 0; d < nd; d++) {
		LSF_REQUIRE(params->field_shape[d] > 0, "field_shape[%d] must be positive, got %d", d, params->field_shape[d]);
		N *= (size_t) params->field_shape[d];
	}
	const size_t pixels = (size_t) rows * cols;
	Arena arena(stream);
	if (memory_kind == LSF_DEVICE)
		return tsdf_generate_device(params, depth_image, rows, cols, camera_pose, image_y_coordinate, nd, field_out, stream);
	unsigned short* depth_dev = nullptr;
	float* field_dev = nullptr;
	LSF_TRY(arena.alloc(&depth_dev, pixels));
	LSF_CUDA(cudaMemcpyAsync(depth_dev, depth_image, pixels * sizeof(unsigned short), cudaMemcpyHostToDevice, stream));
	LSF_TRY(arena.alloc(&field_dev, N));
	LSF_TRY(tsdf_generate_device(params, depth_dev, rows, cols, camera_pose, image_y_coordinate, nd, field_dev, stream));
	return from_device(field_dev, field_out, N, LSF_HOST, stream);
}
