// rigid.cu -- rigid SDF-2-SDF tracker in 2D on the GPU (the step before the non-rigid alignment in
// a full pipeline).
//
// Replaces Sdf2SdfOptimizer2d::optimize, reference cpp/src/rigid_optimization/sdf_2_sdf_optimizer2d.cpp:63-124, with
// gradient_wrt_twist (sdf_gradient_wrt_transformation2d.cpp:18-50), math::gradient 2D (cpp/src/math/gradients.tpp:248-283)
// and transformation_vector_to_matrix2d / 3d (cpp/src/math/transformation.cpp:12-47).
//
// Per iteration: the live TSDF field is generated under the current twist by the TSDF kernels (tsdf.cu), then ONE
// kernel forms every voxel's gradient with respect to the twist and reduces the normal equations A (3 x 3, symmetric),
// b (3) and the energy -- per-thread partial sums in double, a block reduction, and the last block to finish adds the
// block partials in a fixed order (deterministic). The host reads the ten sums (80 bytes), solves the 3 x 3 system in
// float32 the way the reference does (cofactor inverse, twist += rate * (optimal - twist)) and builds the next pose.
// The double-precision sums are more accurate than the reference's sequential float32 accumulation (the 3 x 3 system is
// ill-conditioned: the two differ by ~4e-5 on the twist of the reference's test case). The twist of every iteration agrees
// with the CPU oracle's `double_sums` mode to 2e-6, with its float32-sum mode and with the reference's Python tracker
// within the 1e-4 the reference itself asserts between its C++ and Python trackers (tests/test_sdf_2_sdf_optimizer.py:166).
#include "common.cuh"

#include <cmath>
#include <cstring>

namespace lsf {
namespace {

constexpr int SUMS = 10;  // A00 A01 A02 A11 A12 A22 b0 b1 b2 energy

struct RigidArgs {
	const float* __restrict__ live;
	const float* __restrict__ canonical;
	int H, W;
	int offset_x, offset_z;
	float voxel_size;
	float eta;
	float inverse_twist[6];  // rows 0..1 of transformation_vector_to_matrix2d(-twist)
	float twist[3];
	double* partials;        // [gridDim.x][SUMS]
	double* sums;            // [SUMS] of this iteration
	unsigned* arrivals;      // zero before the launch; reset by the last block
};

__global__ void __launch_bounds__(256) k_sdf2sdf_normal_equations(const RigidArgs a) {
	double acc[SUMS];
#pragma unroll
	for (int k = 0; k < SUMS; k++) acc[k] = 0.0;
	const int N = a.H * a.W;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
		const int y = i / a.W, x = i - y * a.W;
		const float centre = a.live[i];
		// math::gradient, 2D: central differences, one-sided at the borders; component 0 along the columns
		float g0, g1;
		if (x == 0) g0 = a.live[i + 1] - centre;
		else if (x == a.W - 1) g0 = centre - a.live[i - 1];
		else g0 = 0.5f * (a.live[i + 1] - a.live[i - 1]);
		if (y == 0) g1 = a.live[i + a.W] - centre;
		else if (y == a.H - 1) g1 = centre - a.live[i - a.W];
		else g1 = 0.5f * (a.live[i + a.W] - a.live[i - a.W]);
		const float x_voxel = (float) (x + a.offset_x) * a.voxel_size;
		const float z_voxel = (float) (y + a.offset_z) * a.voxel_size;
		const float* M = a.inverse_twist;
		const float t0 = (M[0] * x_voxel + M[1] * z_voxel) + M[2] * 1.f;
		const float t1 = (M[3] * x_voxel + M[4] * z_voxel) + M[5] * 1.f;
		const float g[3] = { (g0 * 1.f + g1 * 0.f) / a.voxel_size, (g0 * 0.f + g1 * 1.f) / a.voxel_size,
				(g0 * t1 + g1 * -t0) / a.voxel_size };
		const float canonical = a.canonical[i];
		const float residual = (canonical - centre) + ((g[0] * a.twist[0] + g[1] * a.twist[1]) + g[2] * a.twist[2]);
		acc[0] += (double) (g[0] * g[0]);
		acc[1] += (double) (g[0] * g[1]);
		acc[2] += (double) (g[0] * g[2]);
		acc[3] += (double) (g[1] * g[1]);
		acc[4] += (double) (g[1] * g[2]);
		acc[5] += (double) (g[2] * g[2]);
		acc[6] += (double) (residual * g[0]);
		acc[7] += (double) (residual * g[1]);
		acc[8] += (double) (residual * g[2]);
		const float cw = canonical <= -a.eta ? 0.f : 1.f, lw = centre <= -a.eta ? 0.f : 1.f;
		const float d = canonical * cw - centre * lw;
		acc[9] += (double) (d * d);
	}
	__shared__ double warp_sums[8][SUMS];
	__shared__ bool last;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int k = 0; k < SUMS; k++) {
		double v = acc[k];
#pragma unroll
		for (int offset = 16; offset > 0; offset >>= 1) v += __shfl_xor_sync(0xffffffffu, v, offset);
		if (lane == 0) warp_sums[warp][k] = v;
	}
	__syncthreads();
	if (threadIdx.x < SUMS) {
		double v = 0.0;
		for (int w = 0; w < 8; w++) v += warp_sums[w][threadIdx.x];
		a.partials[blockIdx.x * SUMS + threadIdx.x] = v;
		__threadfence();
	}
	__syncthreads();
	if (threadIdx.x == 0) last = atomicAdd(a.arrivals, 1u) == gridDim.x - 1;
	__syncthreads();
	if (last && threadIdx.x < SUMS) {
		__threadfence();
		double v = 0.0;
		for (unsigned b = 0; b < gridDim.x; b++) v += a.partials[b * SUMS + threadIdx.x];
		a.sums[threadIdx.x] = v;
		if (threadIdx.x == 0) *a.arrivals = 0;
	}
}

void matrix2d(const float twist[3], float m[9]) {  // transformation.cpp:12-19
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

// transformation.cpp:21-47 for the twist (t0, 0, t1, 0, theta, 0): quaternion about y -> rotation matrix
void pose_of_twist(const float twist[3], float m[16]) {
	float axis = twist[2];
	const float theta = sqrtf((0.f * 0.f + axis * axis) + 0.f * 0.f);
	if (fabsf(theta) > 1e-14) axis /= theta;
	const float w = cosf(theta / 2), y = sinf(theta / 2) * axis;
	const float ty = 2.f * y, twy = ty * w, tyy = ty * y;
	const float R[9] = { 1.f - (tyy + 0.f), 0.f, 0.f + twy, 0.f, 1.f - (0.f + 0.f), 0.f, 0.f - twy, 0.f, 1.f - (0.f + tyy) };
	const float translation[3] = { twist[0], 0.f, twist[1] };
	for (int i = 0; i < 3; i++) {
		for (int j = 0; j < 3; j++) m[4 * i + j] = R[3 * i + j];
		m[4 * i + 3] = translation[i];
	}
	m[12] = m[13] = m[14] = 0.f;
	m[15] = 1.f;
}

void invert3(const float a[9], float inv[9]) {  // cofactors / determinant, like Eigen's fixed-size 3 x 3 inverse
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
}

}  // namespace
}  // namespace lsf

using namespace lsf;

extern "C" int lsf_sdf2sdf_optimize_2d(const lsf_tsdf_params* tsdf_generation_parameters, float rate,
		int maximum_iteration_count, int image_y_coordinate, const float* canonical_field,
		const unsigned short* live_depth_image, int rows, int cols, float eta, const float* initial_camera_pose,
		float* twist_matrix_out, float* twists_out, float* optimal_twists_out, float* energies_out, int memory_kind,
		void* stream_handle) {
	(void) initial_camera_pose;  // the reference accepts and never reads it (sdf_2_sdf_optimizer2d.cpp:63-124)
	cudaStream_t stream = static_cast<cudaStream_t>(stream_handle);
	LSF_REQUIRE(tsdf_generation_parameters && canonical_field && live_depth_image && twist_matrix_out,
			"tsdf_generation_parameters, canonical_field, live_depth_image and twist_matrix_out must not be NULL");
	LSF_REQUIRE(maximum_iteration_count >= 0, "maximum_iteration_count must not be negative");
	const int W = tsdf_generation_parameters->field_shape[0], H = tsdf_generation_parameters->field_shape[1];
	LSF_REQUIRE(W >= 2 && H >= 2, "the field must have at least 2 x 2 voxels, got %d x %d", H, W);
	LSF_REQUIRE(rows > 0 && cols > 0, "depth image must not be empty, got %d x %d", rows, cols);
	const size_t N = (size_t) H * W, pixels = (size_t) rows * cols;
	Arena arena(stream);
	const float* canonical_dev = nullptr;
	LSF_TRY(to_device(arena, canonical_field, N, memory_kind, stream, &canonical_dev));
	const unsigned short* depth_dev = live_depth_image;
	if (memory_kind == LSF_HOST) {
		unsigned short* staged = nullptr;
		LSF_TRY(arena.alloc(&staged, pixels));
		LSF_CUDA(cudaMemcpyAsync(staged, live_depth_image, pixels * sizeof(unsigned short), cudaMemcpyHostToDevice, stream));
		depth_dev = staged;
	}
	const unsigned blocks = std::min<unsigned>(div_up((long long) N, 256), 64u);
	float* live_dev = nullptr;
	double *partials = nullptr, *sums = nullptr;
	unsigned* arrivals = nullptr;
	LSF_TRY(arena.alloc(&live_dev, N));
	LSF_TRY(arena.alloc(&partials, (size_t) blocks * SUMS));
	LSF_TRY(arena.alloc(&sums, SUMS));
	LSF_TRY(arena.alloc(&arrivals, 1));
	LSF_CUDA(cudaMemsetAsync(arrivals, 0, sizeof(unsigned), stream));
	float twist[3] = { 0.f, 0.f, 0.f };
	for (int iteration = 0; iteration < maximum_iteration_count; iteration++) {
		float pose[16];
		pose_of_twist(twist, pose);
		LSF_TRY(tsdf_generate_device(tsdf_generation_parameters, depth_dev, rows, cols, pose, image_y_coordinate, 2, live_dev,
				stream));
		RigidArgs a;
		a.live = live_dev;
		a.canonical = canonical_dev;
		a.H = H;
		a.W = W;
		a.offset_x = tsdf_generation_parameters->array_offset[0];
		a.offset_z = tsdf_generation_parameters->array_offset[1];
		a.voxel_size = tsdf_generation_parameters->voxel_size;
		a.eta = eta;
		const float negated[3] = { -twist[0], -twist[1], -twist[2] };
		float inverse[9];
		matrix2d(negated, inverse);
		std::memcpy(a.inverse_twist, inverse, sizeof(a.inverse_twist));
		std::memcpy(a.twist, twist, sizeof(a.twist));
		a.partials = partials;
		a.sums = sums;
		a.arrivals = arrivals;
		k_sdf2sdf_normal_equations<<<counted(blocks), 256, 0, stream>>>(a);
		LSF_CUDA(cudaGetLastError());
		double host_sums[SUMS];
		LSF_CUDA(cudaMemcpyAsync(host_sums, sums, sizeof(host_sums), cudaMemcpyDeviceToHost, stream));
		LSF_CUDA(cudaStreamSynchronize(stream));
		const float A[9] = { (float) host_sums[0], (float) host_sums[1], (float) host_sums[2], (float) host_sums[1],
				(float) host_sums[3], (float) host_sums[4], (float) host_sums[2], (float) host_sums[4], (float) host_sums[5] };
		const float b[3] = { (float) host_sums[6], (float) host_sums[7], (float) host_sums[8] };
		float inverse_A[9], optimal[3];
		invert3(A, inverse_A);
		for (int r = 0; r < 3; r++)
			optimal[r] = (inverse_A[3 * r] * b[0] + inverse_A[3 * r + 1] * b[1]) + inverse_A[3 * r + 2] * b[2];
		for (int r = 0; r < 3; r++) twist[r] = twist[r] + rate * (optimal[r] - twist[r]);
		for (int r = 0; r < 3; r++) {
			if (twists_out) twists_out[3 * iteration + r] = twist[r];
			if (optimal_twists_out) optimal_twists_out[3 * iteration + r] = optimal[r];
		}
		if (energies_out) energies_out[iteration] = .5f * (float) host_sums[9];
	}
	matrix2d(twist, twist_matrix_out);
	return LSF_OK;
}
