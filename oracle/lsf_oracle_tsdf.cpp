// lsf_oracle_tsdf.cpp -- CPU restatement of the reference's projective TSDF generation from a depth image
// (SURVEY.md 8f row f2). TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline leg, never by the product.
//
// Follows, line by line:
//   3D  Generator<Tensor3>::generate__none      /root/reference/cpp/src/tsdf/generator_tensor.tpp:40-101
//   2D  Generator<MatrixXf>::generate__none     /root/reference/cpp/src/tsdf/generator_matrix.tpp:33-93
//   compute_TSDF_value, is_voxel_out_of_bounds  /root/reference/cpp/src/tsdf/common.hpp:29-51
//   Parameters                                   /root/reference/cpp/src/tsdf/parameters.hpp:31-56
// (Python twins: tsdf/generation.py:356-437 (3D), :142-200 (2D); they clip at depth <= 0 instead of the near clipping
// distance and test the rounded pixel coordinates -- reproduced with near_clipping_distance = 0.)
// float32 arithmetic in the reference's order, products of the 4x4 / 3x3 matrices summed with the column index ascending.
// Pinned by the reference's own goldens cpp/tests/data/test_data_tsdf.hpp:148-191 (test_tsdf.cpp:47-94) and by runs of the
// reference's Python generator (tests/golden/reference_tsdf_runs.npz).
// One deliberate difference: the reference tests voxel_image against [0, cols) x [0, rows) BEFORE rounding
// (common.hpp:43-50 with margin 0), so a voxel that projects to x in [cols - 0.5, cols) is rounded to pixel `cols` and
// read outside the image; here such voxels keep the default value.
#include "lsf_oracle.h"

#include <cmath>

extern "C" int orc_tsdf_generate(const orc_tsdf_params* p, const unsigned short* depth_image, int rows, int cols,
		const float* pose, int image_y_coordinate, int nd, float* field) {
	if (p->filtering_method != 0) return -1;  // only FilteringMethod::NONE is restated
	const int sx = p->field_shape[0], sy = p->field_shape[1], sz = nd == 3 ? p->field_shape[2] : 1;
	const float half_width = (float) (((double) (float) p->narrow_band_width_voxels / 2.) * (double) p->voxel_size);
	const float* P = p->projection_matrix;
	const long long count = (long long) sx * sy * sz;
#pragma omp parallel for
	for (long long i = 0; i < count; i++) field[i] = 1.0f;
	if (nd == 2) {
		// field(y_field, x_field): numpy [y_field][x_field]; x_field counts x, y_field counts depth (z)
#pragma omp parallel for
		for (int y_field = 0; y_field < sy; y_field++)
			for (int x_field = 0; x_field < sx; x_field++) {
				const float x_voxel = (float) (x_field + p->array_offset[0]) * p->voxel_size;
				const float y_voxel = 0.0f;
				const float z_voxel = (float) (y_field + p->array_offset[1]) * p->voxel_size;
				float cam[3];
				for (int r = 0; r < 3; r++)
					cam[r] = ((pose[4 * r] * x_voxel + pose[4 * r + 1] * y_voxel) + pose[4 * r + 2] * z_voxel) + pose[4 * r + 3] * 1.0f;
				if (cam[2] <= p->near_clipping_distance) continue;
				const float image_x = (((P[0] * cam[0] + P[1] * cam[1]) + P[2] * cam[2])) / cam[2];
				const float image_y = (float) image_y_coordinate;
				if (image_x < 0.0f || image_x >= (float) cols || image_y < 0.0f || image_y >= (float) rows) continue;
				const int ix = (int) (image_x + 0.5f);
				if (ix >= cols) continue;  // see header
				const float depth = (float) depth_image[(long long) image_y_coordinate * cols + ix] * p->depth_unit_ratio;
				if (depth <= 0.0f) continue;
				const float sd = depth - cam[2];
				field[(long long) y_field * sx + x_field] = sd < -half_width ? -1.0f : (sd > half_width ? 1.0f : sd / half_width);
			}
		return 0;
	}
#pragma omp parallel for collapse(2)
	for (int x_field = 0; x_field < sx; x_field++)
		for (int y_field = 0; y_field < sy; y_field++)
			for (int z_field = 0; z_field < sz; z_field++) {
				const float x_voxel = (float) (x_field + p->array_offset[0]) * p->voxel_size;
				const float y_voxel = (float) (y_field + p->array_offset[1]) * p->voxel_size;
				const float z_voxel = (float) (z_field + p->array_offset[2]) * p->voxel_size;
				float cam[3];
				for (int r = 0; r < 3; r++)
					cam[r] = ((pose[4 * r] * x_voxel + pose[4 * r + 1] * y_voxel) + pose[4 * r + 2] * z_voxel) + pose[4 * r + 3] * 1.0f;
				if (cam[2] <= p->near_clipping_distance) continue;
				const float image_x = (((P[0] * cam[0] + P[1] * cam[1]) + P[2] * cam[2])) / cam[2];
				const float image_y = (((P[3] * cam[0] + P[4] * cam[1]) + P[5] * cam[2])) / cam[2];
				if (image_x < 0.0f || image_x >= (float) cols || image_y < 0.0f || image_y >= (float) rows) continue;
				const int ix = (int) (image_x + 0.5f), iy = (int) (image_y + 0.5f);
				if (ix >= cols || iy >= rows) continue;  // see header
				const float depth = (float) depth_image[(long long) iy * cols + ix] * p->depth_unit_ratio;
				if (depth <= 0.0f) continue;
				const float sd = depth - cam[2];
				field[((long long) x_field * sy + y_field) * sz + z_field] =
						sd < -half_width ? -1.0f : (sd > half_width ? 1.0f : sd / half_width);
			}
	return 0;
}
