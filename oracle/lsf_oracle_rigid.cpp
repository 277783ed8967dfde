// lsf_oracle_rigid.cpp -- CPU restatement of the reference's rigid SDF-2-SDF tracker in 2D.
// TEST INFRASTRUCTURE ONLY: imported by tests/, never by the product.
//
// Follows, line by line:
//   Sdf2SdfOptimizer2d::optimize            /root/reference/cpp/src/rigid_optimization/sdf_2_sdf_optimizer2d.cpp:63-124
//   gradient_wrt_twist                      /root/reference/cpp/src/rigid_optimization/sdf_gradient_wrt_transformation2d.cpp:18-50
//   transformation_vector_to_matrix2d / 3d  /root/reference/cpp/src/math/transformation.cpp:12-47
//       (quaternion -> rotation matrix as Eigen::Quaternionf::toRotationMatrix evaluates it)
//   math::gradient (2D)                     /root/reference/cpp/src/math/gradients.tpp:248-283   (orc_gradient2d)
//   tsdf::Generator2d::generate             orc_tsdf_generate (lsf_oracle_tsdf.cpp)
// (Python twin: rigid_opt/sdf_2_sdf_optimizer2d.py:62-137 with rigid_opt/sdf_gradient_field.py.)
// float32 accumulation of the normal equations in the reference's loop order (columns outer, rows inner); the 3 x 3
// inverse by cofactors like Eigen's fixed-size inverse. `double_sums` != 0 keeps every per-voxel term in float32 but adds
// them up in double (what the GPU reduction does; the order of a double sum of <= 2^20 float32 terms does not reach the
// float32 result) -- the tests use it to compare the device path tightly, and the float32 mode to bound the distance
// between the two (the 3 x 3 system is ill-conditioned: ~4e-5 on the twist for the reference's test case). The reference holds no golden for its C++ tracker; it asserts
// C++ == Python within 1e-4 on the twist matrix (tests/test_sdf_2_sdf_optimizer.py:81-166). Pinned the same way: by a run
// of the reference's Python tracker made in the build container (tests/golden/reference_rigid.npz), at 1e-4.
// Like the reference, `initial_camera_pose` is accepted by the boundary and ignored (sdf_2_sdf_optimizer2d.cpp:63-124 never
// reads it).
#include "lsf_oracle.h"

#include <cmath>
#include <vector>

namespace {

void matrix2d(const float twist[3], float m[9]) {  // transformation.cpp:12-19: cos / sin of a double, stored as float
	const double theta = twist[2];
	m[0] = (float) cos(theta);
	m[1] = (float) -sin(theta);
	m[2] = twist[0];
	m[3] = (float) sin(theta);
	m[4] = (float) cos(theta);
	m[5] = twist[1];
	m[6] = 0.f;
	m[7] = 0.f;
	m[8] = 1.f;
}

void matrix3d(const float twist6[6], float m[16]) {  // transformation.cpp:21-47
	float rotation[3] = { twist6[3], twist6[4], twist6[5] };
	const float theta = sqrtf((rotation[0] * rotation[0] + rotation[1] * rotation[1]) + rotation[2] * rotation[2]);
	if (fabsf(theta) > 1e-14)
		for (int i = 0; i < 3; i++) rotation[i] /= theta;
	const float w = cosf(theta / 2), s = sinf(theta / 2);
	const float x = s * rotation[0], y = s * rotation[1], z = s * rotation[2];
	// Eigen::QuaternionBase::toRotationMatrix
	const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
	const float twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y,
			tzz = tz * z;
	const float R[9] = { 1.f - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1.f - (txx + tzz), tyz - twx, txz - twy, tyz + twx,
			1.f - (txx + tyy) };
	for (int i = 0; i < 3; i++) {
		for (int j = 0; j < 3; j++) m[4 * i + j] = R[3 * i + j];
		m[4 * i + 3] = twist6[i];
	}
	m[12] = m[13] = m[14] = 0.f;
	m[15] = 1.f;
}

bool invert3(const float a[9], float inv[9]) {  // cofactors / determinant (Eigen compute_inverse_size3_helper)
	const float c00 = a[4] * a[8] - a[5] * a[7], c10 = a[5] * a[6] - a[3] * a[8], c20 = a[3] * a[7] - a[4] * a[6];
	const float determinant = (c00 * a[0] + c10 * a[1]) + c20 * a[2];
	const float inverse_determinant = 1.f / determinant;
	inv[0] = c00 * inverse_determinant;
	inv[3] = c10 * inverse_determinant;
	inv[6] = c20 * inverse_determinant;
	inv[1] = (a[2] * a[7] - a[1] * a[8]) * inverse_determinant;
	inv[4] = (a[0] * a[8] - a[2] * a[6]) * inverse_determinant;
	inv[7] = (a[1] * a[6] - a[0] * a[7]) * inverse_determinant;
	inv[2] = (a[1] * a[5] - a[2] * a[4]) * inverse_determinant;
	inv[5] = (a[2] * a[3] - a[0] * a[5]) * inverse_determinant;
	inv[8] = (a[0] * a[4] - a[1] * a[3]) * inverse_determinant;
	return std::isfinite(inverse_determinant);
}

}  // namespace

extern "C" int orc_sdf2sdf_optimize(const orc_tsdf_params* tsdf_parameters, float rate, int maximum_iteration_count,
		int image_y_coordinate, const float* canonical_field, const unsigned short* live_depth_image, int rows, int cols,
		float eta, int double_sums, float* twist_matrix_out, float* twists_out, float* energies_out) {
	const int W = tsdf_parameters->field_shape[0], H = tsdf_parameters->field_shape[1];
	const float voxel_size = tsdf_parameters->voxel_size;
	std::vector<float> live((size_t) H * W), gradient((size_t) H * W * 2);
	float twist[3] = { 0.f, 0.f, 0.f };
	for (int iteration = 0; iteration < maximum_iteration_count; iteration++) {
		float A[9] = { 0 }, b[3] = { 0 };
		double A_double[9] = { 0 }, b_double[3] = { 0 }, energy_double = 0.0;  // double_sums: float32 terms, double sums
		const float twist6[6] = { twist[0], 0.f, twist[1], 0.f, twist[2], 0.f };
		float pose[16];
		matrix3d(twist6, pose);
		const int status = orc_tsdf_generate(tsdf_parameters, live_depth_image, rows, cols, pose, image_y_coordinate, 2, live.data());
		if (status != 0) return status;
		orc_gradient2d(live.data(), H, W, gradient.data());
		const float negated[3] = { -twist[0], -twist[1], -twist[2] };
		float inverse_twist_matrix[9];
		matrix2d(negated, inverse_twist_matrix);
		const float* M = inverse_twist_matrix;
		for (int x_field = 0; x_field < W; x_field++)
			for (int y_field = 0; y_field < H; y_field++) {
				const float x_voxel = (float) (x_field + tsdf_parameters->array_offset[0]) * voxel_size;
				const float z_voxel = (float) (y_field + tsdf_parameters->array_offset[1]) * voxel_size;
				const float t0 = (M[0] * x_voxel + M[1] * z_voxel) + M[2] * 1.f;
				const float t1 = (M[3] * x_voxel + M[4] * z_voxel) + M[5] * 1.f;
				const size_t i = (size_t) y_field * W + x_field;
				const float g0 = gradient[2 * i], g1 = gradient[2 * i + 1];
				const float g[3] = { (g0 * 1.f + g1 * 0.f) / voxel_size, (g0 * 0.f + g1 * 1.f) / voxel_size,
						(g0 * t1 + g1 * -t0) / voxel_size };
				for (int r = 0; r < 3; r++)
					for (int c = 0; c < 3; c++) {
						A[3 * r + c] += g[r] * g[c];
						A_double[3 * r + c] += (double) (g[r] * g[c]);
					}
				const float residual = (canonical_field[i] - live[i]) + ((g[0] * twist[0] + g[1] * twist[1]) + g[2] * twist[2]);
				for (int r = 0; r < 3; r++) {
					b[r] += residual * g[r];
					b_double[r] += (double) (residual * g[r]);
				}
			}
		float energy = 0.f;
		for (size_t i = 0; i < (size_t) H * W; i++) {
			const float cw = canonical_field[i] <= -eta ? 0.f : 1.f, lw = live[i] <= -eta ? 0.f : 1.f;
			const float d = canonical_field[i] * cw - live[i] * lw;
			energy += d * d;
			energy_double += (double) (d * d);
		}
		energy *= .5f;
		if (double_sums) {
			for (int k = 0; k < 9; k++) A[k] = (float) A_double[k];
			for (int k = 0; k < 3; k++) b[k] = (float) b_double[k];
			energy = .5f * (float) energy_double;
		}
		float inverse[9];
		invert3(A, inverse);
		float optimal[3];
		for (int r = 0; r < 3; r++) optimal[r] = (inverse[3 * r] * b[0] + inverse[3 * r + 1] * b[1]) + inverse[3 * r + 2] * b[2];
		for (int r = 0; r < 3; r++) twist[r] = twist[r] + rate * (optimal[r] - twist[r]);
		if (twists_out)
			for (int r = 0; r < 3; r++) twists_out[3 * iteration + r] = twist[r];
		if (energies_out) energies_out[iteration] = energy;
	}
	matrix2d(twist, twist_matrix_out);
	return 0;
}
