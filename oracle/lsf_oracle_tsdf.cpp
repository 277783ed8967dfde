// lsf_oracle_tsdf.cpp -- CPU restatement of the reference's projective TSDF generation from a depth image
// (SURVEY.md 8f row f2). TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline leg, never by the product.
//
// Follows, line by line:
//   3D  Generator<Tensor3>::generate__none      /root/reference/cpp/src/tsdf/generator_tensor.tpp:40-101
//       Generator<Tensor3>::generate__ewa_aux   /root/reference/cpp/src/tsdf/generator_tensor.tpp:126-240
//   2D  Generator<MatrixXf>::generate__none     /root/reference/cpp/src/tsdf/generator_matrix.tpp:33-93
//       Generator<MatrixXf>::generate__ewa_aux  /root/reference/cpp/src/tsdf/generator_matrix.tpp:118-207
//   compute_TSDF_value, is_voxel_out_of_bounds  /root/reference/cpp/src/tsdf/common.hpp:29-51
//   compute_covariance_camera_space, compute_sampling_bounds[_inclusive], compute_voxel_EWA_image_space,
//   compute_voxel_EWA_voxel_space[_inclusive]   /root/reference/cpp/src/tsdf/ewa_common.hpp:32-235
//   compute_centered_ellipse_bound_points       /root/reference/cpp/src/math/conics.cpp:44-59
//   Parameters, FilteringMethod                 /root/reference/cpp/src/tsdf/parameters.hpp:31-56, interpolation_method.hpp:39-46
// (Python twins: tsdf/generation.py:356-437 (3D), :130-217 (2D), tsdf/ewa.py:59-185,230-600.)
// float32 arithmetic in the reference's order; products of the small fixed-size matrices are summed with the inner index
// ascending (Eigen's own evaluation order for these sizes can differ in the last bit: the reference's tests compare at
// 1e-6, and so do the tests of this restatement).
// Pinned by the reference's own goldens cpp/tests/data/test_data_tsdf.hpp (test_tsdf.cpp:47-333),
// tests/test_data/ewa_test_data.py (tests/test_tsdf_ewa.py) and by runs of the reference's Python generators
// (tests/golden/reference_tsdf.npz, made by tests/golden/make_tsdf_golden.py).
// Reference quirks kept: the 2D EWA generators clip against the GLOBAL constant near_clipping_distance = 0.05 instead of
// the parameter (generator_matrix.tpp:146); the x / y bounds of a tilted ellipse are exchanged (conics.cpp:54-56).
// One deliberate difference: filtering NONE tests voxel_image against [0, cols) x [0, rows) BEFORE rounding
// (common.hpp:43-50 with margin 0), so a voxel that projects to x in [cols - 0.5, cols) is rounded to pixel `cols` and
// read outside the image; here such voxels keep the default value. Non-square 2D fields index out of bounds in the
// reference (generator_matrix.tpp:55-56 decodes with x_size for both axes); here they are generated with the evident
// meaning.
#include "lsf_oracle.h"

#include <cfloat>
#include <cmath>

namespace {

inline float tsdf_value(float sd, float half_width) {  // common.hpp:32-40
	return sd < -half_width ? -1.0f : (sd > half_width ? 1.0f : sd / half_width);
}

struct Context {
	const orc_tsdf_params* p;
	const unsigned short* depth;
	int rows, cols;
	const float* pose;
	float half_width;
	float covariance[3][3];  // compute_covariance_camera_space
	float threshold;         // squared_radius_threshold
};

// the value of one voxel; false: the voxel keeps the default value 1
inline bool voxel_value(const Context& c, int nd, float x_voxel, float y_voxel, float z_voxel, int image_y_coordinate,
		float* out) {
	const orc_tsdf_params* p = c.p;
	const float* pose = c.pose;
	const float* P = p->projection_matrix;
	float cam[3];
	for (int r = 0; r < 3; r++)
		cam[r] = ((pose[4 * r] * x_voxel + pose[4 * r + 1] * y_voxel) + pose[4 * r + 2] * z_voxel) + pose[4 * r + 3] * 1.0f;
	const int method = p->filtering_method;
	const float near = (nd == 2 && method != 0) ? 0.05f : p->near_clipping_distance;
	if (cam[2] <= near) return false;
	const float image_x = (((P[0] * cam[0] + P[1] * cam[1]) + P[2] * cam[2])) / cam[2];
	const float image_y = nd == 2 ? (float) image_y_coordinate : (((P[3] * cam[0] + P[4] * cam[1]) + P[5] * cam[2])) / cam[2];
	const int rows = c.rows, cols = c.cols;
	if (method == 0) {
		if (image_x < 0.0f || image_x >= (float) cols || image_y < 0.0f || image_y >= (float) rows) return false;
		// 3D rounds in double (`int(voxel_image(0) + 0.5)`, generator_tensor.tpp:86-87), 2D in float (generator_matrix.tpp:77)
		const int ix = nd == 2 ? (int) (image_x + 0.5f) : (int) ((double) image_x + 0.5);
		const int iy = nd == 2 ? image_y_coordinate : (int) ((double) image_y + 0.5);
		if (ix >= cols || iy >= rows) return false;  // see header
		const float depth = (float) c.depth[(long long) iy * cols + ix] * p->depth_unit_ratio;
		if (depth <= 0.0f) return false;
		*out = tsdf_value(depth - cam[2], c.half_width);
		return true;
	}
	// ---- EWA: generate__ewa_aux
	if (image_x < -3.0f || image_x >= (float) (cols + 3) || image_y < -3.0f || image_y >= (float) (rows + 3)) return false;
	const float ray_distance = sqrtf((cam[0] * cam[0] + cam[1] * cam[1]) + cam[2] * cam[2]);
	const float z_cam_squared = cam[2] * cam[2];
	const float inv_z_cam = 1.0f / cam[2];
	const float J[3][3] = { { inv_z_cam, 0.0f, -cam[0] / z_cam_squared }, { 0.0f, inv_z_cam, -cam[1] / z_cam_squared },
			{ cam[0] / ray_distance, cam[1] / ray_distance, cam[2] / ray_distance } };
	float T[3][3], R2[2][2];
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++)
			T[i][j] = (J[i][0] * c.covariance[0][j] + J[i][1] * c.covariance[1][j]) + J[i][2] * c.covariance[2][j];
	for (int i = 0; i < 2; i++)
		for (int j = 0; j < 2; j++) R2[i][j] = (T[i][0] * J[j][0] + T[i][1] * J[j][1]) + T[i][2] * J[j][2];
	const float S[2][2] = { { P[0], P[1] }, { P[3], P[4] } };
	float SR[2][2], F2[2][2];
	for (int i = 0; i < 2; i++)
		for (int j = 0; j < 2; j++) SR[i][j] = S[i][0] * R2[0][j] + S[i][1] * R2[1][j];
	for (int i = 0; i < 2; i++)
		for (int j = 0; j < 2; j++) F2[i][j] = (SR[i][0] * S[j][0] + SR[i][1] * S[j][1]) + (i == j ? 1.0f : 0.0f);
	const float determinant = F2[0][0] * F2[1][1] - F2[1][0] * F2[0][1];
	const float inverse_determinant = 1.0f / determinant;
	const float Q[2][2] = { { F2[1][1] * inverse_determinant, -F2[0][1] * inverse_determinant },
			{ -F2[1][0] * inverse_determinant, F2[0][0] * inverse_determinant } };
	// compute_centered_ellipse_bound_points
	const float A = Q[0][0], B = Q[0][1] * 2.0f, C = Q[1][1], F = c.threshold;
	float bound_x, bound_y;
	if (fabsf(B) < FLT_EPSILON) {
		bound_x = sqrtf(F / A);
		bound_y = sqrtf(F / C);
	} else {
		const float B_squared = B * B;
		bound_x = sqrtf(F / (C - B_squared / (4.0f * A)));
		bound_y = sqrtf(F / (A - B_squared / (4.0f * C)));
	}
	int x_start = (int) (image_x - bound_x);
	int x_end = (int) ceilf(image_x + bound_x + 1.0f);
	int y_start = (int) (image_y - bound_y);
	int y_end = (int) ceilf(image_y + bound_y + 1.0f);
	if (x_start >= cols || x_end <= 0 || y_start >= rows || y_end <= 0) return false;
	if (method != 5) {
		x_start = x_start > 0 ? x_start : 0;
		x_end = x_end < cols ? x_end : cols;
		y_start = y_start > 0 ? y_start : 0;
		y_end = y_end < rows ? y_end : rows;
	}
	float weights_sum = 0.0f, value_sum = 0.0f;
	for (int x_sample = x_start; x_sample < x_end; x_sample++)
		for (int y_sample = y_start; y_sample < y_end; y_sample++) {
			const float sx = (float) x_sample - image_x, sy = (float) y_sample - image_y;
			const float dist_sq = (sx * Q[0][0] + sy * Q[1][0]) * sx + (sx * Q[0][1] + sy * Q[1][1]) * sy;
			if (dist_sq > c.threshold) continue;
			const float weight = expf(-0.5f * dist_sq);
			if (method == 5 && (y_sample < 0 || y_sample >= rows || x_sample < 0 || x_sample >= cols)) {
				value_sum += weight;
				weights_sum += weight;
				continue;
			}
			const float surface_depth = (float) c.depth[(long long) y_sample * cols + x_sample] * p->depth_unit_ratio;
			if (surface_depth <= 0.0f) continue;
			if (method == 3)
				value_sum += weight * surface_depth;
			else
				value_sum += weight * tsdf_value(surface_depth - cam[2], c.half_width);
			weights_sum += weight;
		}
	if (method == 3) {
		if (value_sum <= 0.0f) {
			*out = 1.0f;
			return true;
		}
		*out = tsdf_value(value_sum / weights_sum - cam[2], c.half_width);
		return true;
	}
	*out = weights_sum == 0.0f ? 1.0f : value_sum / weights_sum;
	return true;
}

}  // namespace

extern "C" int orc_tsdf_generate(const orc_tsdf_params* p, const unsigned short* depth_image, int rows, int cols,
		const float* pose, int image_y_coordinate, int nd, float* field) {
	const int method = p->filtering_method;
	if (method != 0 && method != 3 && method != 4 && method != 5) return -1;  // bilinear methods: "Not yet implemented"
	const int sx = p->field_shape[0], sy = p->field_shape[1], sz = nd == 3 ? p->field_shape[2] : 1;
	Context c;
	c.p = p;
	c.depth = depth_image;
	c.rows = rows;
	c.cols = cols;
	c.pose = pose;
	c.half_width = (float) (((double) (float) p->narrow_band_width_voxels / 2.) * (double) p->voxel_size);
	c.threshold = 4.0f * p->voxel_size * p->smoothing_factor;
	{
		const float s = p->voxel_size * p->smoothing_factor;
		float M[3][3];
		for (int i = 0; i < 3; i++)
			for (int j = 0; j < 3; j++) M[i][j] = pose[4 * i + j] * s;  // R * (I * s)
		for (int i = 0; i < 3; i++)
			for (int j = 0; j < 3; j++)
				c.covariance[i][j] = (M[i][0] * pose[4 * j] + M[i][1] * pose[4 * j + 1]) + M[i][2] * pose[4 * j + 2];
	}
	const long long count = (long long) sx * sy * sz;
#pragma omp parallel for
	for (long long i = 0; i < count; i++) field[i] = 1.0f;
	if (nd == 2) {
		// field(y_field, x_field): numpy [y_field][x_field]; x_field counts x, y_field counts depth (z)
#pragma omp parallel for schedule(dynamic, 1)
		for (int y_field = 0; y_field < sy; y_field++)
			for (int x_field = 0; x_field < sx; x_field++) {
				const float x_voxel = (float) (x_field + p->array_offset[0]) * p->voxel_size;
				const float z_voxel = (float) (y_field + p->array_offset[1]) * p->voxel_size;
				float value;
				if (voxel_value(c, 2, x_voxel, 0.0f, z_voxel, image_y_coordinate, &value))
					field[(long long) y_field * sx + x_field] = value;
			}
		return 0;
	}
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
	for (int x_field = 0; x_field < sx; x_field++)
		for (int y_field = 0; y_field < sy; y_field++)
			for (int z_field = 0; z_field < sz; z_field++) {
				const float x_voxel = (float) (x_field + p->array_offset[0]) * p->voxel_size;
				const float y_voxel = (float) (y_field + p->array_offset[1]) * p->voxel_size;
				const float z_voxel = (float) (z_field + p->array_offset[2]) * p->voxel_size;
				float value;
				if (voxel_value(c, 3, x_voxel, y_voxel, z_voxel, 0, &value))
					field[((long long) x_field * sy + y_field) * sz + z_field] = value;
			}
	return 0;
}
