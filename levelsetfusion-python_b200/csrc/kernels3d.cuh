// kernels3d.cuh -- sm_100a kernels of the 3D warp-field optimisation path.
//
// Device-side data layout (DESIGN.md "Data layout in HBM"):
//   * scalar fields           float  f[X][Y][Z]                      (Z contiguous)
//   * vector fields (warp, g) three planes  p[c][X][Y][Z]            (SoA; AoS [X][Y][Z][3] only at the API)
//   * live level "pack"       float4 {live, dlive/dx, dlive/dy, dlive/dz} on a grid padded by 2 voxels on
//                             every side and pre-filled with the out-of-bounds constants {1,0,0,0}; one
//                             LDG.128 per trilinear tap fetches the TSDF value and its gradient, and the
//                             padding turns the reference's per-tap bounds tests into two integer clamps.
//
// All arithmetic follows the reference's operation order in float32 without FMA (compiled with
// --fmad=false); citations give the reference file:line each kernel restates.
#pragma once

#include "common.cuh"

namespace lsf {

struct Grid3 {
	int X, Y, Z;
	long long N;
	__host__ __device__ Grid3() : X(0), Y(0), Z(0), N(0) {}
	__host__ __device__ Grid3(int x, int y, int z) : X(x), Y(y), Z(z), N((long long) x * y * z) {}
	__host__ __device__ Grid3 half() const {
		return Grid3(X / 2, Y / 2, Z / 2);
	}
	__host__ __device__ Grid3 twice() const {
		return Grid3(X * 2, Y * 2, Z * 2);
	}
	// padded pack geometry
	__host__ __device__ int PY() const {
		return Y + 4;
	}
	__host__ __device__ int PZ() const {
		return Z + 4;
	}
	__host__ __device__ long long padded_count() const {
		return (long long) (X + 4) * (Y + 4) * (Z + 4);
	}
	__host__ __device__ long long padded_index(int x, int y, int z) const {
		return ((long long) (x + 2) * PY() + (y + 2)) * PZ() + (z + 2);
	}
};

#ifdef __CUDACC__

constexpr int BLOCK_Z = 32;
constexpr int BLOCK_Y = 8;

inline dim3 block3() {
	return dim3(BLOCK_Z, BLOCK_Y, 1);
}
inline dim3 grid3(const Grid3& g) {
	return dim3(div_up(g.Z, BLOCK_Z), div_up(g.Y, BLOCK_Y), g.X);
}

#define LSF_VOXEL_3D(g)                                        \
	const int z = blockIdx.x * BLOCK_Z + threadIdx.x;          \
	const int y = blockIdx.y * BLOCK_Y + threadIdx.y;          \
	const int x = blockIdx.z;                                  \
	const bool in_grid = (z < (g).Z) && (y < (g).Y);           \
	const long long idx = ((long long) x * (g).Y + y) * (g).Z + z

// ---------------------------------------------------------------------------------------------- float4 arithmetic
__device__ __forceinline__ float4 operator+(float4 a, float4 b) {
	return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 operator*(float4 a, float s) {
	return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
__device__ __forceinline__ float4 operator*(float s, float4 a) {
	return make_float4(s * a.x, s * a.y, s * a.z, s * a.w);
}
__device__ __forceinline__ float4 operator/(float4 a, float s) {
	return make_float4(a.x / s, a.y / s, a.z / s, a.w / s);
}

// ---------------------------------------------------------------------------------------------- stencil terms
// central difference, one-sided at the borders: reference cpp/src/math/gradients.tpp:447-494
__device__ __forceinline__ float central_difference(const float* __restrict__ f, long long idx, long long stride,
		int i, int n) {
	if (n < 2) return 0.0f;
	if (i == 0) return f[idx + stride] - f[idx];
	if (i == n - 1) return f[idx] - f[idx - stride];
	return 0.5f * (f[idx + stride] - f[idx - stride]);
}

// one axis of the replicated-border Laplacian: reference cpp/src/math/gradients.tpp:28-35,114-171
__device__ __forceinline__ float laplace_term(const float* __restrict__ f, long long idx, long long stride, int i,
		int n) {
	if (n < 2) return 0.0f;
	if (i == 0) return f[idx + stride] - f[idx];
	if (i == n - 1) return f[idx - stride] - f[idx];
	return (f[idx + stride] - 2.0f * f[idx]) + f[idx - stride];
}

__device__ __forceinline__ float laplacian_at(const float* __restrict__ f, long long idx, int x, int y, int z,
		const Grid3& g) {
	float acc = laplace_term(f, idx, (long long) g.Y * g.Z, x, g.X);
	acc += laplace_term(f, idx, g.Z, y, g.Y);
	acc += laplace_term(f, idx, 1, z, g.Z);
	return acc;
}

// ---------------------------------------------------------------------------------------------- trilinear gather
// reference cpp/src/nonrigid_optimization/field_warping.tpp:80-134: lookup = index + warp component,
// base = floor, ratio = lookup - base, interpolation along z, then y, then x. Out-of-bounds taps read the
// pack's padding ({1,0,0,0} for the optimizer: TSDF -> 1, gradient -> 0).
__device__ __forceinline__ float4 gather4(const float4* __restrict__ pack, const Grid3& g, int x, int y, int z,
		float wx, float wy, float wz) {
	const float lookup_x = (float) x + wx;
	const float lookup_y = (float) y + wy;
	const float lookup_z = (float) z + wz;
	int bx = __float2int_rd(lookup_x);
	int by = __float2int_rd(lookup_y);
	int bz = __float2int_rd(lookup_z);
	const float rx = lookup_x - (float) bx, ry = lookup_y - (float) by, rz = lookup_z - (float) bz;
	const float ix = 1.0f - rx, iy = 1.0f - ry, iz = 1.0f - rz;
	bx = min(max(bx, -2), g.X);
	by = min(max(by, -2), g.Y);
	bz = min(max(bz, -2), g.Z);
	const long long sy = g.PZ(), sx = (long long) g.PY() * g.PZ();
	const float4* p = pack + g.padded_index(bx, by, bz);
	const float4 v000 = __ldg(p), v001 = __ldg(p + 1);
	const float4 v010 = __ldg(p + sy), v011 = __ldg(p + sy + 1);
	const float4 v100 = __ldg(p + sx), v101 = __ldg(p + sx + 1);
	const float4 v110 = __ldg(p + sx + sy), v111 = __ldg(p + sx + sy + 1);
	const float4 i00 = v000 * iz + v001 * rz;
	const float4 i01 = v010 * iz + v011 * rz;
	const float4 i10 = v100 * iz + v101 * rz;
	const float4 i11 = v110 * iz + v111 * rz;
	const float4 i0 = i00 * iy + i01 * ry;
	const float4 i1 = i10 * iy + i11 * ry;
	return i0 * ix + i1 * rx;
}

// ---------------------------------------------------------------------------------------------- pack construction
static __global__ void k_fill4(float4* __restrict__ p, long long n, float4 value) {
	const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) p[i] = value;
}

// finest level: gradient of the live field (reference math::gradient, gradients.tpp:438-495) packed with it
static __global__ void k_gradient_pack3d(const float* __restrict__ live, float4* __restrict__ pack, Grid3 g) {
	LSF_VOXEL_3D(g);
	if (!in_grid) return;
	const float gx = central_difference(live, idx, (long long) g.Y * g.Z, x, g.X);
	const float gy = central_difference(live, idx, g.Z, y, g.Y);
	const float gz = central_difference(live, idx, 1, z, g.Z);
	pack[g.padded_index(x, y, z)] = make_float4(live[idx], gx, gy, gz);
}

// plain scalar or AoS vector field -> pack (primitives and parity tests): channels 1 -> .x, 3 -> .y.z.w
static __global__ void k_pack_field3d(const float* __restrict__ field, int channels, float4* __restrict__ pack, Grid3 g) {
	LSF_VOXEL_3D(g);
	if (!in_grid) return;
	float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
	if (channels == 1) {
		v.x = field[idx];
	} else {
		v.y = field[idx * 3];
		v.z = field[idx * 3 + 1];
		v.w = field[idx * 3 + 2];
	}
	pack[g.padded_index(x, y, z)] = v;
}

static __global__ void k_gradient3d(const float* __restrict__ field, float* __restrict__ out_aos, Grid3 g) {
	LSF_VOXEL_3D(g);
	if (!in_grid) return;
	out_aos[idx * 3 + 0] = central_difference(field, idx, (long long) g.Y * g.Z, x, g.X);
	out_aos[idx * 3 + 1] = central_difference(field, idx, g.Z, y, g.Y);
	out_aos[idx * 3 + 2] = central_difference(field, idx, 1, z, g.Z);
}

// ---------------------------------------------------------------------------------------------- layout changes
static __global__ void k_aos_to_planes(const float* __restrict__ aos, float* __restrict__ planes, long long n, int channels) {
	const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	for (int c = 0; c < channels; c++) planes[c * n + i] = aos[i * channels + c];
}

static __global__ void k_planes_to_aos(const float* __restrict__ planes, float* __restrict__ aos, long long n, int channels) {
	const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	for (int c = 0; c < channels; c++) aos[i * channels + c] = planes[c * n + i];
}

// ---------------------------------------------------------------------------------------------- gather (primitive)
// reference warp / warp_with_replacement, field_warping.tpp:212-225
static __global__ void k_gather_pack3d(const float4* __restrict__ pack, const float* __restrict__ warp_aos,
		float* __restrict__ out, int channels, Grid3 g) {
	LSF_VOXEL_3D(g);
	if (!in_grid) return;
	const float4 v = gather4(pack, g, x, y, z, warp_aos[idx * 3], warp_aos[idx * 3 + 1], warp_aos[idx * 3 + 2]);
	if (channels == 1) {
		out[idx] = v.x;
	} else {
		out[idx * 3] = v.y;
		out[idx * 3 + 1] = v.z;
		out[idx * 3 + 2] = v.w;
	}
}

// ---------------------------------------------------------------------------------------------- iteration, stage 1
// reference Optimizer::optimize_iteration, cpp/src/nonrigid_optimization/hierarchical/optimizer.tpp:186-200:
//   resampled_live = warp(live, w); resampled_grad = warp_with_replacement(grad, w, 0)
//   diff = resampled_live - canonical; data_gradient = resampled_grad * diff
//   g = data_gradient * amplifier [ - laplacian(g_prev) * strength ]
// FUSE_UPDATE (no Sobolev kernel configured) additionally performs :207-211 in the same pass:
//   w -= g * rate; max ||g||^2 -> max_sq_bits[iteration]
struct HierIterArgs {
	const float4* pack;      // padded live level
	const float* canonical;  // [N]
	const float* warp;       // planes, read
	float* warp_out;         // planes, written when FUSE_UPDATE (may alias warp: point-wise)
	const float* g_prev;     // planes, read when TIKHONOV (must not alias g_out)
	float* g_out;            // planes, may be nullptr when nothing reads it later
	Grid3 g;
	float amplifier, strength, rate, threshold;
	unsigned* max_sq_bits;
	int iteration;
	int check_convergence;
	// slab decomposition along axis 0 (slab.py): the fields are an allocation of g.X planes (owned planes + halo
	// planes); only planes [x_begin, x_end) are processed. Allocation plane 0 is global plane x_origin of a level of
	// X_global planes; the pack covers global planes [pack_origin, pack_origin + pack_X). Whole volume: x_begin 0,
	// x_end X, x_origin 0, X_global X, pack_origin 0, pack_X X, no interior pack edges.
	int x_begin, x_end, x_origin, X_global;
	int pack_X, pack_origin, pack_interior_low, pack_interior_high;
	int* violation;          // set to 1 when a gather would leave the rank's pack region (nullptr: whole volume)
	// batch of independent pairs (lsf_hier_optimize_3d_batch; the reference's multi-pair loop,
	// run_hierarchical_optimizer3d_multipair.py:403-406): the plane fields hold the pairs one after the other along
	// axis 0 (g.X = pairs * batch_X planes, component stride g.N), every pair has its own padded pack
	// (pack + pair * batch_pack_stride) and its own row of convergence slots (max_sq_bits + pair * batch_slot_stride).
	// Border rules, gather clamps and the filter's zero padding apply per pair. batch_X == 0: one volume.
	int batch_X;
	int batch_chunks;        // x-chunks per pair (stage-1 kernels: blockIdx.z = pair * batch_chunks + chunk)
	long long batch_pack_stride;
	int batch_slot_stride;
};

inline void whole_volume(HierIterArgs& a) {
	a.x_begin = 0;
	a.x_end = a.g.X;
	a.x_origin = 0;
	a.X_global = a.g.X;
	a.pack_X = a.g.X;
	a.pack_origin = 0;
	a.pack_interior_low = a.pack_interior_high = 0;
	a.violation = nullptr;
	a.batch_X = 0;
	a.batch_chunks = 1;
	a.batch_pack_stride = 0;
	a.batch_slot_stride = 0;
}

template<bool TIKHONOV, bool FUSE_UPDATE>
static __global__ void __launch_bounds__(BLOCK_Z * BLOCK_Y) k_hier_gradient3d(HierIterArgs a) {
	if (a.check_convergence && level_converged(a.max_sq_bits, a.iteration, a.threshold)) return;
	const Grid3 g = a.g;
	LSF_VOXEL_3D(g);
	float sq = 0.0f;
	if (in_grid) {
		const float wx = a.warp[idx], wy = a.warp[g.N + idx], wz = a.warp[2 * g.N + idx];
		const float4 s = gather4(a.pack, g, x, y, z, wx, wy, wz);
		const float diff = s.x - a.canonical[idx];
		float gx = (s.y * diff) * a.amplifier;
		float gy = (s.z * diff) * a.amplifier;
		float gz = (s.w * diff) * a.amplifier;
		if (TIKHONOV) {
			gx = gx - laplacian_at(a.g_prev, idx, x, y, z, g) * a.strength;
			gy = gy - laplacian_at(a.g_prev + g.N, idx, x, y, z, g) * a.strength;
			gz = gz - laplacian_at(a.g_prev + 2 * g.N, idx, x, y, z, g) * a.strength;
		}
		if (a.g_out != nullptr) {
			a.g_out[idx] = gx;
			a.g_out[g.N + idx] = gy;
			a.g_out[2 * g.N + idx] = gz;
		}
		if (FUSE_UPDATE) {
			a.warp_out[idx] = wx - gx * a.rate;
			a.warp_out[g.N + idx] = wy - gy * a.rate;
			a.warp_out[2 * g.N + idx] = wz - gz * a.rate;
			sq = 0.0f + gx * gx;
			sq += gy * gy;
			sq += gz * gz;
		}
	}
	if (FUSE_UPDATE) block_atomic_max(sq, a.max_sq_bits + a.iteration);
}

// ---------------------------------------------------------------------------------------------- iteration, stage 2
// One pass of the separable filter along AXIS (reference convolve_with_kernel, convolution.cpp:221-332:
// zero padded, flipped taps, accumulation from 0.0f over taps i-r..i+r ascending). The axis-2 pass is the
// last one (reference order x, y, z) and, when FINAL, also applies optimizer.tpp:207-211.
struct ConvArgs {
	const float* in;   // planes
	float* out;        // planes (must not alias in)
	float* warp;       // planes, updated in place when FINAL
	Grid3 g;
	Taps taps;
	float rate, threshold;
	unsigned* max_sq_bits;
	int iteration;
	int check_convergence;
	int channels;
};

template<int AXIS, bool FINAL>
static __global__ void __launch_bounds__(BLOCK_Z * BLOCK_Y) k_convolve_axis3d(ConvArgs a) {
	if (a.check_convergence && level_converged(a.max_sq_bits, a.iteration, a.threshold)) return;
	const Grid3 g = a.g;
	LSF_VOXEL_3D(g);
	float sq = 0.0f;
	if (in_grid) {
		const int i = AXIS == 0 ? x : (AXIS == 1 ? y : z);
		const int n = AXIS == 0 ? g.X : (AXIS == 1 ? g.Y : g.Z);
		const long long stride = AXIS == 0 ? (long long) g.Y * g.Z : (AXIS == 1 ? g.Z : 1);
		const int r = a.taps.radius;
		for (int c = 0; c < a.channels; c++) {
			const float* line = a.in + c * g.N + idx;
			float acc = 0.0f;
			for (int j = 0; j < a.taps.size; j++) {
				const int src = i - r + j;
				const float value = (src >= 0 && src < n) ? __ldg(line + (long long) (j - r) * stride) : 0.0f;
				acc += value * a.taps.k[j];
			}
			a.out[c * g.N + idx] = acc;
			if (FINAL) {
				a.warp[c * g.N + idx] = a.warp[c * g.N + idx] - acc * a.rate;
				sq += acc * acc;
			}
		}
	}
	if (FINAL) block_atomic_max(sq, a.max_sq_bits + a.iteration);
}

// max ||v||^2 of a planes field (reference locate_max_norm, statistics.tpp:76-100)
static __global__ void k_max_sq_norm_planes(const float* __restrict__ planes, long long n, int channels,
		unsigned* __restrict__ max_sq_bits) {
	float best = 0.0f;
	for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n;
			i += (long long) gridDim.x * blockDim.x) {
		float sq = 0.0f;
		for (int c = 0; c < channels; c++) sq += planes[c * n + i] * planes[c * n + i];
		if (sq > best) best = sq;
	}
	block_atomic_max(best, max_sq_bits);
}

static __global__ void k_laplacian_planes3d(const float* __restrict__ in, float* __restrict__ out, int channels, Grid3 g) {
	LSF_VOXEL_3D(g);
	if (!in_grid) return;
	for (int c = 0; c < channels; c++) out[c * g.N + idx] = laplacian_at(in + c * g.N, idx, x, y, z, g);
}

// ---------------------------------------------------------------------------------------------- restrict x2
// Element accessors let the same kernel serve plain scalar fields, planes and the padded float4 pack.
struct PlainAccess {
	const float* src;
	float* dst;
	__device__ float load(const Grid3& g, int x, int y, int z) const {
		return src[((long long) x * g.Y + y) * g.Z + z];
	}
	__device__ void store(const Grid3& g, int x, int y, int z, float v) const {
		dst[((long long) x * g.Y + y) * g.Z + z] = v;
	}
};
struct PackAccess {
	const float4* src;
	float4* dst;
	__device__ float4 load(const Grid3& g, int x, int y, int z) const {
		return src[g.padded_index(x, y, z)];
	}
	__device__ void store(const Grid3& g, int x, int y, int z, float4 v) const {
		dst[g.padded_index(x, y, z)] = v;
	}
};

// AVERAGE: reference downsampleX2_average, cpp/src/math/resampling.tpp:385-417 (first index fastest, /8)
template<typename Access>
static __global__ void k_downsample_average3d(Access acc, Grid3 src, Grid3 dst) {
	LSF_VOXEL_3D(dst);
	(void) idx;
	if (!in_grid) return;
	const int sx = 2 * x, sy = 2 * y, sz = 2 * z;
	auto sum = acc.load(src, sx, sy, sz) + acc.load(src, sx + 1, sy, sz);
	sum = sum + acc.load(src, sx, sy + 1, sz);
	sum = sum + acc.load(src, sx + 1, sy + 1, sz);
	sum = sum + acc.load(src, sx, sy, sz + 1);
	sum = sum + acc.load(src, sx + 1, sy, sz + 1);
	sum = sum + acc.load(src, sx, sy + 1, sz + 1);
	sum = sum + acc.load(src, sx + 1, sy + 1, sz + 1);
	acc.store(dst, x, y, z, sum / 8.0f);
}

// LINEAR: reference downsampleX2_linear, cpp/src/math/resampling.tpp:544-656: 4^3 tent on the replicate-padded
// field; the four weight groups are summed in the reference's tap order (table below, offsets from 2t).
static __constant__ signed char c_linear_taps3d[64][3] = {
		// group 0 (8 taps, weight 27/512)
		{ 0, 0, 0 }, { 1, 0, 0 }, { 0, 1, 0 }, { 1, 1, 0 }, { 0, 0, 1 }, { 1, 0, 1 }, { 0, 1, 1 }, { 1, 1, 1 },
		// group 1 (24 taps, weight 9/512)
		{ -1, 0, 0 }, { 0, -1, 0 }, { 0, 0, -1 }, { 2, 0, 0 }, { 1, -1, 0 }, { 1, 0, -1 }, { -1, 1, 0 }, { 0, 2, 0 },
		{ 0, 1, -1 }, { 2, 1, 0 }, { 1, 2, 0 }, { 1, 1, -1 }, { -1, 0, 1 }, { 0, -1, 1 }, { 0, 0, 2 }, { 2, 0, 1 },
		{ 1, -1, 1 }, { 1, 0, 2 }, { -1, 1, 1 }, { 0, 2, 1 }, { 0, 1, 2 }, { 2, 1, 1 }, { 1, 2, 1 }, { 1, 1, 2 },
		// group 2 (24 taps, weight 3/512)
		{ -1, -1, 0 }, { 0, -1, -1 }, { -1, 0, -1 }, { 2, -1, 0 }, { 1, -1, -1 }, { 2, 0, -1 }, { -1, 2, 0 },
		{ 0, 2, -1 }, { -1, 1, -1 }, { 2, 2, 0 }, { 1, 2, -1 }, { 2, 1, -1 }, { -1, -1, 1 }, { 0, -1, 2 }, { -1, 0, 2 },
		{ 2, -1, 1 }, { 1, -1, 2 }, { 2, 0, 2 }, { -1, 2, 1 }, { 0, 2, 2 }, { -1, 1, 2 }, { 2, 2, 1 }, { 1, 2, 2 },
		{ 2, 1, 2 },
		// group 3 (8 taps, weight 1/512)
		{ -1, -1, -1 }, { 2, -1, -1 }, { -1, 2, -1 }, { 2, 2, -1 }, { -1, -1, 2 }, { 2, -1, 2 }, { -1, 2, 2 },
		{ 2, 2, 2 } };

template<typename Access>
static __global__ void k_downsample_linear3d(Access acc, Grid3 src, Grid3 dst) {
	LSF_VOXEL_3D(dst);
	(void) idx;
	if (!in_grid) return;
	const float c0 = 0.052734375f * 4.0f, c1 = 0.017578125f * 4.0f, c2 = 0.005859375f * 4.0f,
			c3 = 0.001953125f * 4.0f;
	auto tap = [&](int t) {
		const int tx = min(max(2 * x + c_linear_taps3d[t][0], 0), src.X - 1);
		const int ty = min(max(2 * y + c_linear_taps3d[t][1], 0), src.Y - 1);
		const int tz = min(max(2 * z + c_linear_taps3d[t][2], 0), src.Z - 1);
		return acc.load(src, tx, ty, tz);
	};
	auto s0 = tap(0);
	for (int t = 1; t < 8; t++) s0 = s0 + tap(t);
	auto s1 = tap(8);
	for (int t = 9; t < 32; t++) s1 = s1 + tap(t);
	auto s2 = tap(32);
	for (int t = 33; t < 56; t++) s2 = s2 + tap(t);
	auto s3 = tap(56);
	for (int t = 57; t < 64; t++) s3 = s3 + tap(t);
	acc.store(dst, x, y, z, (((c0 * s0 + c1 * s1) + c2 * s2) + c3 * s3) * 0.25f);
}

// ---------------------------------------------------------------------------------------------- prolong x2
// NEAREST: reference upsampleX2_nearest, cpp/src/math/resampling.tpp:103-126
static __global__ void k_upsample_nearest3d(const float* __restrict__ src, float* __restrict__ dst, int channels,
		Grid3 sg, Grid3 dg) {
	LSF_VOXEL_3D(dg);
	if (!in_grid) return;
	const long long sidx = ((long long) (x >> 1) * sg.Y + (y >> 1)) * sg.Z + (z >> 1);
	for (int c = 0; c < channels; c++) dst[c * dg.N + idx] = src[c * sg.N + sidx];
}

// 2D linear prolongation of a strided source matrix evaluated at one target element,
// reference upsampleX2_linear (matrix), cpp/src/math/resampling.tpp:130-216.
__device__ __forceinline__ float upsample_linear2d_at(const float* __restrict__ s, long long row_stride,
		long long col_stride, int H, int W, int r, int c) {
	auto S = [&](int rr, int cc) {return s[rr * row_stride + cc * col_stride];};
	const int UH = 2 * H, UW = 2 * W;
	if (r == 0 || r == UH - 1) {
		const int sr = r == 0 ? 0 : H - 1;
		if (c == 0) return S(sr, 0);
		if (c == UW - 1) return S(sr, W - 1);
		const int sc = (c + 1) >> 1;  // "current" source column of the pair
		const float prev = S(sr, sc - 1), cur = S(sr, sc);
		return (c & 1) ? 0.75f * prev + 0.25f * cur : 0.25f * prev + 0.75f * cur;
	}
	if (c == 0 || c == UW - 1) {
		const int sc = c == 0 ? 0 : W - 1;
		const int sr = (r + 1) >> 1;
		const float prev = S(sr - 1, sc), cur = S(sr, sc);
		return (r & 1) ? 0.75f * prev + 0.25f * cur : 0.25f * prev + 0.75f * cur;
	}
	const int sr = (r - 1) >> 1, sc = (c - 1) >> 1;
	const float v00 = S(sr, sc), v01 = S(sr, sc + 1), v10 = S(sr + 1, sc), v11 = S(sr + 1, sc + 1);
	if (r & 1) {
		if (c & 1) return ((0.5625f * v00 + 0.1875f * v01) + 0.1875f * v10) + 0.0625f * v11;
		return ((0.1875f * v00 + 0.5625f * v01) + 0.0625f * v10) + 0.1875f * v11;
	}
	if (c & 1) return ((0.1875f * v00 + 0.0625f * v01) + 0.5625f * v10) + 0.1875f * v11;
	return ((0.0625f * v00 + 0.1875f * v01) + 0.1875f * v10) + 0.5625f * v11;
}

// LINEAR 3D: reference upsampleX2_linear (tensor), cpp/src/math/resampling.tpp:218-322. Faces are 2D
// prolongations of the source faces, written in the order near/far x, near/far y, near/far z, so on shared
// edges the z faces win, then y; the interior interpolates along x, then y, then z.
static __global__ void k_upsample_linear3d(const float* __restrict__ src, float* __restrict__ dst, int channels, Grid3 sg,
		Grid3 dg) {
	LSF_VOXEL_3D(dg);
	if (!in_grid) return;
	const long long sX = (long long) sg.Y * sg.Z, sY = sg.Z, sZ = 1;
	for (int c = 0; c < channels; c++) {
		const float* s = src + c * sg.N;
		float value;
		if (z == 0 || z == dg.Z - 1) {
			value = upsample_linear2d_at(s + (z == 0 ? 0 : (sg.Z - 1) * sZ), sX, sY, sg.X, sg.Y, x, y);
		} else if (y == 0 || y == dg.Y - 1) {
			value = upsample_linear2d_at(s + (y == 0 ? 0 : (sg.Y - 1) * sY), sX, sZ, sg.X, sg.Z, x, z);
		} else if (x == 0 || x == dg.X - 1) {
			value = upsample_linear2d_at(s + (x == 0 ? 0 : (sg.X - 1) * sX), sY, sZ, sg.Y, sg.Z, y, z);
		} else {
			const int xs = (x - 1) >> 1, ys = (y - 1) >> 1, zs = (z - 1) >> 1;
			const float* p = s + xs * sX + ys * sY + zs * sZ;
			const float ax = (x & 1) ? 0.75f : 0.25f, bx = (x & 1) ? 0.25f : 0.75f;
			const float ay = (y & 1) ? 0.75f : 0.25f, by = (y & 1) ? 0.25f : 0.75f;
			const float az = (z & 1) ? 0.75f : 0.25f, bz = (z & 1) ? 0.25f : 0.75f;
			const float x00 = ax * p[0] + bx * p[sX];
			const float x10 = ax * p[sY] + bx * p[sX + sY];
			const float x01 = ax * p[sZ] + bx * p[sX + sZ];
			const float x11 = ax * p[sY + sZ] + bx * p[sX + sY + sZ];
			const float y0 = ay * x00 + by * x10;
			const float y1 = ay * x01 + by * x11;
			value = az * y0 + bz * y1;
		}
		dst[c * dg.N + idx] = value;
	}
}

#endif  // __CUDACC__

}  // namespace lsf
