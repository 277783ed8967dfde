// hier_telemetry.cuh -- per-iteration telemetry of the hierarchical optimizers (2D and 3D): the fields and statistics
// the reference's OptimizerWithTelemetry collects around every optimize_iteration call
// (cpp/src/nonrigid_optimization/hierarchical/optimizer_with_telemetry.tpp:83-182):
//   * OptimizationIterationData (telemetry/optimization_iteration_data.tpp): live pyramid level, warp field after the
//     iteration, data-term gradient (resampled live gradient * diff, before the amplifier) and Tikhonov-term gradient
//     (Laplacian of the previous gradient, before the strength) of every iteration;
//   * the numbers behind the verbosity flags: max update length, mean and standard deviation of diff = warped live -
//     canonical (math/statistics.tpp:255-271), normalised data energy 1e6 * mean(diff^2), normalised Tikhonov energy
//     1e6 * 0.5 * mean((sum of the entries of the Jacobian of the previous gradient)^2) (tpp:139-147).
// This is a diagnostic path (the reference copies whole fields per iteration too): it runs one iteration at a time with
// straightforward kernels next to the production iteration kernels, which it leaves untouched. The reductions run on
// the GPU with double accumulators (compared with a tolerance, like the reference's almost_equal).
#pragma once

#include "common.cuh"

#include <cmath>
#include <vector>

namespace lsf {

#ifdef __CUDACC__

// out[0] += sum_i (x[i] - shift), out[1] += sum_i (x[i] - shift)^2
static __global__ void k_telemetry_moments(const float* __restrict__ x, long long n, float shift, double* __restrict__ out) {
	double s1 = 0.0, s2 = 0.0;
	for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
		const float v = x[i] - shift;
		s1 += (double) v;
		s2 += (double) v * (double) v;
	}
	__shared__ double part[2][8];
#pragma unroll
	for (int offset = 16; offset > 0; offset >>= 1) {
		s1 += __shfl_xor_sync(0xffffffffu, s1, offset);
		s2 += __shfl_xor_sync(0xffffffffu, s2, offset);
	}
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (lane == 0) {
		part[0][warp] = s1;
		part[1][warp] = s2;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		double t1 = 0.0, t2 = 0.0;
		for (int w = 0; w < (int) (blockDim.x >> 5); w++) {
			t1 += part[0][w];
			t2 += part[1][w];
		}
		atomicAdd(out, t1);
		atomicAdd(out + 1, t2);
	}
}

// Jacobian of a D-component plane field f[c][N] on a grid n0 x n1 x n2 (unused trailing axes have extent 1): central
// differences, one-sided on the faces (reference math::gradient of a vector field, gradients.tpp:286-387);
// out[0] += (sum of all D x D entries)^2 per voxel
static __global__ void k_telemetry_jacobian_energy(const float* __restrict__ f, int D, long long N, int n0, int n1, int n2,
		double* __restrict__ out) {
	double total = 0.0;
	const int n[3] = { n0, n1, n2 };
	const long long stride[3] = { (long long) n1 * n2, n2, 1 };
	for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long) gridDim.x * blockDim.x) {
		const int p[3] = { (int) (i / stride[0]), (int) ((i / stride[1]) % n1), (int) (i % n2) };
		float sum = 0.0f;
		for (int axis = 0; axis < 3; axis++) {
			if (n[axis] < 2) continue;
			for (int c = 0; c < D; c++) {
				const float* v = f + c * N + i;
				float d;
				if (p[axis] == 0) d = v[stride[axis]] - v[0];
				else if (p[axis] == n[axis] - 1) d = v[0] - v[-stride[axis]];
				else d = 0.5f * (v[stride[axis]] - v[-stride[axis]]);
				sum += d;
			}
		}
		total += (double) sum * (double) sum;
	}
#pragma unroll
	for (int offset = 16; offset > 0; offset >>= 1) total += __shfl_xor_sync(0xffffffffu, total, offset);
	__shared__ double part[8];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (lane == 0) part[warp] = total;
	__syncthreads();
	if (threadIdx.x == 0) {
		double t = 0.0;
		for (int w = 0; w < (int) (blockDim.x >> 5); w++) t += part[w];
		atomicAdd(out, t);
	}
}

static __global__ void k_telemetry_planes_to_aos(const float* __restrict__ planes, float* __restrict__ aos, long long n,
		int channels) {
	const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	for (int c = 0; c < channels; c++) aos[i * channels + c] = planes[c * n + i];
}

#endif  // __CUDACC__

// Scratch of the telemetry path for one optimize() call (sized for the finest level).
struct TelemetryScratch {
	float* diff = nullptr;          // [N]
	float* data_planes = nullptr;   // [D][N] data-term gradient
	float* tikhonov_planes = nullptr;
	float* aos = nullptr;           // [N][D] conversion buffer
	float* live_level = nullptr;    // [N]
	double* sums = nullptr;         // 4 doubles on the device
	std::vector<float> host_live, host_warp, host_data, host_tikhonov;
	int allocate(Arena& arena, size_t N, int D, bool tikhonov, bool want_fields) {
		LSF_TRY(arena.alloc(&diff, N));
		LSF_TRY(arena.alloc(&data_planes, N * D));
		if (tikhonov) LSF_TRY(arena.alloc(&tikhonov_planes, N * D));
		LSF_TRY(arena.alloc(&aos, N * D));
		LSF_TRY(arena.alloc(&live_level, N));
		LSF_TRY(arena.alloc(&sums, (size_t) 4));
		if (want_fields) {
			host_live.resize(N);
			host_warp.resize(N * D);
			host_data.resize(N * D);
			if (tikhonov) host_tikhonov.resize(N * D);
		}
		return LSF_OK;
	}
};

#ifdef __CUDACC__

// Statistics of one iteration from `diff` (this iteration) and the previous gradient planes; fills the record's
// mean / std / energies. Synchronises the stream (diagnostic path).
inline int telemetry_statistics(TelemetryScratch& t, long long N, int D, const int* dims, const float* g_prev_planes,
		bool tikhonov, lsf_iteration_record* record, cudaStream_t stream) {
	const unsigned blocks = (unsigned) std::min<long long>(1184, (N + 255) / 256);
	double host[4] = { 0, 0, 0, 0 };
	LSF_CUDA(cudaMemsetAsync(t.sums, 0, 4 * sizeof(double), stream));
	k_telemetry_moments<<<counted(blocks), 256, 0, stream>>>(t.diff, N, 0.0f, t.sums);
	if (tikhonov && g_prev_planes != nullptr)
		k_telemetry_jacobian_energy<<<counted(blocks), 256, 0, stream>>>(g_prev_planes, D, N, dims[0], dims[1],
				D == 3 ? dims[2] : 1, t.sums + 2);
	LSF_CUDA(cudaMemcpyAsync(host, t.sums, 4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
	LSF_CUDA(cudaStreamSynchronize(stream));
	const float mean = (float) (host[0] / (double) N);
	record->mean_tsdf_difference = mean;
	record->normalized_data_energy = (float) (1000000.0 * (host[1] / (double) N));
	record->normalized_tikhonov_energy = tikhonov ? (float) (1000000.0 * 0.5 * (host[2] / (double) N)) : 0.0f;
	// second pass around the mean, like the reference's math::std
	LSF_CUDA(cudaMemsetAsync(t.sums, 0, 2 * sizeof(double), stream));
	k_telemetry_moments<<<counted(blocks), 256, 0, stream>>>(t.diff, N, mean, t.sums);
	LSF_CUDA(cudaMemcpyAsync(host, t.sums, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
	LSF_CUDA(cudaStreamSynchronize(stream));
	record->std_tsdf_difference = (float) std::sqrt(host[1] / (double) N);
	return LSF_OK;
}

// Copies the iteration's fields to the host staging buffers (interleaved [voxel][component]) and fills the record's
// pointers. `warp_planes` = the warp after the iteration.
inline int telemetry_fields(TelemetryScratch& t, long long N, int D, const float* warp_planes, bool tikhonov,
		lsf_iteration_record* record, cudaStream_t stream) {
	const unsigned blocks = (unsigned) ((N + 255) / 256);
	auto fetch = [&](const float* planes, std::vector<float>& host) -> int {
		k_telemetry_planes_to_aos<<<counted(blocks), 256, 0, stream>>>(planes, t.aos, N, D);
		LSF_CUDA(cudaMemcpyAsync(host.data(), t.aos, (size_t) N * D * sizeof(float), cudaMemcpyDeviceToHost, stream));
		LSF_CUDA(cudaStreamSynchronize(stream));
		return LSF_OK;
	};
	LSF_TRY(fetch(warp_planes, t.host_warp));
	LSF_TRY(fetch(t.data_planes, t.host_data));
	if (tikhonov) LSF_TRY(fetch(t.tikhonov_planes, t.host_tikhonov));
	LSF_CUDA(cudaMemcpyAsync(t.host_live.data(), t.live_level, (size_t) N * sizeof(float), cudaMemcpyDeviceToHost, stream));
	LSF_CUDA(cudaStreamSynchronize(stream));
	record->live_field = t.host_live.data();
	record->warp_field = t.host_warp.data();
	record->data_term_gradient = t.host_data.data();
	record->tikhonov_term_gradient = tikhonov ? t.host_tikhonov.data() : nullptr;
	return LSF_OK;
}

#endif  // __CUDACC__

}  // namespace lsf
