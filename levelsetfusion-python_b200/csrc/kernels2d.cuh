// kernels2d.cuh -- sm_100a kernels of the 2D warp-field optimisation path (hierarchical + slavcheva).
//
// Layout mirrors the 3D path: scalar f[H][W] (W contiguous), vector fields as two planes p[c][H][W]
// (component 0 = u displaces along columns, 1 = v along rows: reference field_warping.tpp:159-168),
// live level packed as float4 {live, dlive/dcol, dlive/drow, 0} on a grid padded by 2 and pre-filled with
// the out-of-bounds constants. Arithmetic order follows the reference exactly (no FMA).
#pragma once

#include "common.cuh"
#include "kernels3d.cuh"  // float4 operators, stencil terms, upsample_linear2d_at, block config

namespace lsf {

struct Grid2 {
	int H, W;
	long long N;
	__host__ __device__ Grid2() : H(0), W(0), N(0) {}
	__host__ __device__ Grid2(int h, int w) : H(h), W(w), N((long long) h * w) {}
	__host__ __device__ Grid2 half() const {
		return Grid2(H / 2, W / 2);
	}
	__host__ __device__ int PW() const {
		return W + 4;
	}
	__host__ __device__ long long padded_count() const {
		return (long long) (H + 4) * (W + 4);
	}
	__host__ __device__ long long padded_index(int r, int c) const {
		return (long long) (r + 2) * PW() + (c + 2);
	}
};

#ifdef __CUDACC__

inline dim3 grid2(const Grid2& g) {
	return dim3(div_up(g.W, BLOCK_Z), div_up(g.H, BLOCK_Y), 1);
}

#define LSF_PIXEL_2D(g)                                        \
	const int col = blockIdx.x * BLOCK_Z + threadIdx.x;        \
	const int row = blockIdx.y * BLOCK_Y + threadIdx.y;        \
	const bool in_grid = (col < (g).W) && (row < (g).H);       \
	const long long idx = (long long) row * (g).W + col

// a tap of the gather (element `index` of the padded live pack, in padded row `padded_row`): a plain load; the single-cluster
// level kernel of small fields (hier2d_persistent.cu) keeps the pack in distributed shared memory and defines its own look-up
#ifndef HIER_PACK_TAP
#define HIER_PACK_TAP(pack, padded_row, index) __ldg((pack) + (index))
#endif

// bilinear gather: reference field_warping.tpp:159-189 -- interpolation along y (rows) first, then x
__device__ __forceinline__ float4 gather4_2d(const float4* __restrict__ pack, const Grid2& g, int row, int col,
		float u, float v) {
	const float lookup_x = (float) col + u;
	const float lookup_y = (float) row + v;
	int bx = __float2int_rd(lookup_x);
	int by = __float2int_rd(lookup_y);
	const float rx = lookup_x - (float) bx, ry = lookup_y - (float) by;
	const float ix = 1.0f - rx, iy = 1.0f - ry;
	bx = min(max(bx, -2), g.W);
	by = min(max(by, -2), g.H);
	const long long at = g.padded_index(by, bx);
	const float4 v00 = HIER_PACK_TAP(pack, by + 2, at), v01 = HIER_PACK_TAP(pack, by + 3, at + g.PW());  // (x, y), (x, y+1)
	const float4 v10 = HIER_PACK_TAP(pack, by + 2, at + 1), v11 = HIER_PACK_TAP(pack, by + 3, at + g.PW() + 1);  // (x+1, ..)
	const float4 i0 = v00 * iy + v01 * ry;
	const float4 i1 = v10 * iy + v11 * ry;
	return i0 * ix + i1 * rx;
}

// reference math::gradient (matrix), gradients.tpp:248-283: .x = d/dcol, .y = d/drow
static __global__ void k_gradient_pack2d(const float* __restrict__ live, float4* __restrict__ pack, Grid2 g) {
	LSF_PIXEL_2D(g);
	if (!in_grid) return;
	const float gx = central_difference(live, idx, 1, col, g.W);
	const float gy = central_difference(live, idx, g.W, row, g.H);
	pack[g.padded_index(row, col)] = make_float4(live[idx], gx, gy, 0.0f);
}

static __global__ void k_gradient2d(const float* __restrict__ field, float* __restrict__ out_aos, Grid2 g) {
	LSF_PIXEL_2D(g);
	if (!in_grid) return;
	out_aos[idx * 2 + 0] = central_difference(field, idx, 1, col, g.W);
	out_aos[idx * 2 + 1] = central_difference(field, idx, g.W, row, g.H);
}

static __global__ void k_pack_field2d(const float* __restrict__ field, int channels, float4* __restrict__ pack, Grid2 g) {
	LSF_PIXEL_2D(g);
	if (!in_grid) return;
	float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
	if (channels == 1) {
		v.x = field[idx];
	} else {
		v.y = field[idx * 2];
		v.z = field[idx * 2 + 1];
	}
	pack[g.padded_index(row, col)] = v;
}

static __global__ void k_gather_pack2d(const float4* __restrict__ pack, const float* __restrict__ warp_aos,
		float* __restrict__ out, int channels, Grid2 g) {
	LSF_PIXEL_2D(g);
	if (!in_grid) return;
	const float4 v = gather4_2d(pack, g, row, col, warp_aos[idx * 2], warp_aos[idx * 2 + 1]);
	if (channels == 1) {
		out[idx] = v.x;
	} else {
		out[idx * 2] = v.y;
		out[idx * 2 + 1] = v.z;
	}
}

// replicated-border Laplacian: rows term, then columns term added (reference gradients.tpp:62-101)
__device__ __forceinline__ float laplacian2d_at(const float* __restrict__ f, long long idx, int row, int col,
		const Grid2& g) {
	float acc = laplace_term(f, idx, g.W, row, g.H);
	acc += laplace_term(f, idx, 1, col, g.W);
	return acc;
}

static __global__ void k_laplacian_planes2d(const float* __restrict__ in, float* __restrict__ out, int channels, Grid2 g) {
	LSF_PIXEL_2D(g);
	if (!in_grid) return;
	for (int c = 0; c < channels; c++) out[c * g.N + idx] = laplacian2d_at(in + c * g.N, idx, row, col, g);
}

// ---------------------------------------------------------------------------------------------- hierarchical iteration
struct HierIterArgs2 {
	const float4* pack;
	const float* canonical;
	const float* warp;
	float* warp_out;
	const float* g_prev;
	float* g_out;
	Grid2 g;
	float amplifier, strength, rate, threshold;
	unsigned* max_sq_bits;
	int iteration;
	int check_convergence;
};

// reference optimizer.tpp:186-211 instantiated for MatrixXf / MatrixXv2f (see k_hier_gradient3d)
// one pixel of the gradient stage (body of k_hier_gradient2d; also called by the single-launch level kernel of small
// fields, hier2d_persistent.cu); sq = the pixel's ||g||^2 when FUSE_UPDATE
template<bool TIKHONOV, bool FUSE_UPDATE>
__device__ __forceinline__ void hier_gradient2d_at(const HierIterArgs2& a, int row, int col, long long idx, float& sq) {
	const Grid2& g = a.g;
	const float u = a.warp[idx], v = a.warp[g.N + idx];
	const float4 s = gather4_2d(a.pack, g, row, col, u, v);
	const float diff = s.x - a.canonical[idx];
	float gx = (s.y * diff) * a.amplifier;
	float gy = (s.z * diff) * a.amplifier;
	if (TIKHONOV) {
		gx = gx - laplacian2d_at(a.g_prev, idx, row, col, g) * a.strength;
		gy = gy - laplacian2d_at(a.g_prev + g.N, idx, row, col, g) * a.strength;
	}
	if (a.g_out != nullptr) {
		a.g_out[idx] = gx;
		a.g_out[g.N + idx] = gy;
	}
	if (FUSE_UPDATE) {
		a.warp_out[idx] = u - gx * a.rate;
		a.warp_out[g.N + idx] = v - gy * a.rate;
		sq = 0.0f + gx * gx;
		sq += gy * gy;
	}
}

template<bool TIKHONOV, bool FUSE_UPDATE>
static __global__ void __launch_bounds__(BLOCK_Z * BLOCK_Y) k_hier_gradient2d(HierIterArgs2 a) {
	if (a.check_convergence && level_converged(a.max_sq_bits, a.iteration, a.threshold)) return;
	const Grid2 g = a.g;
	LSF_PIXEL_2D(g);
	float sq = 0.0f;
	if (in_grid) hier_gradient2d_at<TIKHONOV, FUSE_UPDATE>(a, row, col, idx, sq);
	if (FUSE_UPDATE) block_atomic_max(sq, a.max_sq_bits + a.iteration);
}

// ---------------------------------------------------------------------------------------------- separable filter
// reference convolve_with_kernel (matrix), convolution.cpp:147-219: pass 1 runs along the row index
// (axis 0, "y"), pass 2 along the column index; PRESERVE_ZEROS (:23-47,:69-145) writes a zero vector where
// the pass-input vector is exactly zero.
struct ConvArgs2 {
	const float* in;
	float* out;
	float* warp;  // FINAL of the hierarchical optimizer: warp -= out * rate
	Grid2 g;
	Taps taps;
	float rate, threshold;
	unsigned* max_sq_bits;
	int iteration;
	int check_convergence;
	int channels;
	int preserve_zeros;
};

// one pixel of a filter pass (body of k_convolve_axis2d); sq accumulates the pixel's ||out||^2 when FINAL.
// R > 0: the caller knows the filter radius (a.taps.radius == R): the tap loop unrolls (the taps are then read at
// compile-time offsets of the argument block instead of from a stack copy of it) and pixels at least R from both ends of
// their line skip the bounds tests (same taps, same order)
template<int AXIS, bool FINAL, int R = 0>
__device__ __forceinline__ void convolve_axis2d_at(const ConvArgs2& a, int row, int col, long long idx, float& sq) {
	const Grid2& g = a.g;
	const int i = AXIS == 0 ? row : col;
	const int n = AXIS == 0 ? g.H : g.W;
	const long long stride = AXIS == 0 ? g.W : 1;
	const int r = a.taps.radius;
	bool keep_zero = false;
	if (a.preserve_zeros) {
		keep_zero = true;
		for (int c = 0; c < a.channels; c++) keep_zero = keep_zero && (a.in[c * g.N + idx] == 0.0f);
	}
	for (int c = 0; c < a.channels; c++) {
		float acc = 0.0f;
		if (!keep_zero) {
			const float* line = a.in + c * g.N + idx;
			if (R > 0 && i >= R && i + R < n) {
#pragma unroll
				for (int j = 0; j < 2 * R + 1; j++) acc += __ldg(line + (long long) (j - R) * stride) * a.taps.k[j];
			} else if (R > 0) {
#pragma unroll
				for (int j = 0; j < 2 * R + 1; j++) {
					const int src = i - R + j;
					const float value = (src >= 0 && src < n) ? __ldg(line + (long long) (j - R) * stride) : 0.0f;
					acc += value * a.taps.k[j];
				}
			} else {
				for (int j = 0; j < a.taps.size; j++) {
					const int src = i - r + j;
					const float value = (src >= 0 && src < n) ? __ldg(line + (long long) (j - r) * stride) : 0.0f;
					acc += value * a.taps.k[j];
				}
			}
		}
		a.out[c * g.N + idx] = acc;
		if (FINAL) {
			a.warp[c * g.N + idx] = a.warp[c * g.N + idx] - acc * a.rate;
			sq += acc * acc;
		}
	}
}

template<int AXIS, bool FINAL>
static __global__ void __launch_bounds__(BLOCK_Z * BLOCK_Y) k_convolve_axis2d(ConvArgs2 a) {
	if (a.check_convergence && level_converged(a.max_sq_bits, a.iteration, a.threshold)) return;
	const Grid2 g = a.g;
	LSF_PIXEL_2D(g);
	float sq = 0.0f;
	if (in_grid) convolve_axis2d_at<AXIS, FINAL>(a, row, col, idx, sq);
	if (FINAL) block_atomic_max(sq, a.max_sq_bits + a.iteration);
}

// Small fields: all iterations of a level in one cooperative launch (hier2d_persistent.cu). `gradient` / `filter` are the
// arguments of iteration `first_iteration` (gradient.g_prev = the level's g_post, gradient.g_out = its scratch field with a
// Sobolev kernel or the Tikhonov term, else nullptr); the kernel rotates the two fields like enqueue_iteration does.
long long hier2d_persistent_capacity();
int launch_hier2d_persistent(const HierIterArgs2& gradient, const ConvArgs2& filter, bool tikhonov, bool use_kernel,
		float* g_post, float* scratch, int first_iteration, int count, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------- restrict x2
struct PlainAccess2 {
	const float* src;
	float* dst;
	__device__ float load(const Grid2& g, int r, int c) const {
		return src[(long long) r * g.W + c];
	}
	__device__ void store(const Grid2& g, int r, int c, float v) const {
		dst[(long long) r * g.W + c] = v;
	}
};
struct PackAccess2 {
	const float4* src;
	float4* dst;
	__device__ float4 load(const Grid2& g, int r, int c) const {
		return src[g.padded_index(r, c)];
	}
	__device__ void store(const Grid2& g, int r, int c, float4 v) const {
		dst[g.padded_index(r, c)] = v;
	}
};

// AVERAGE: reference downsampleX2_average (matrix), resampling.tpp:358-382
template<typename Access>
static __global__ void k_downsample_average2d(Access acc, Grid2 src, Grid2 dst) {
	LSF_PIXEL_2D(dst);
	(void) idx;
	if (!in_grid) return;
	const int sr = 2 * row, sc = 2 * col;
	auto sum = acc.load(src, sr, sc) + acc.load(src, sr, sc + 1);
	sum = sum + acc.load(src, sr + 1, sc);
	sum = sum + acc.load(src, sr + 1, sc + 1);
	acc.store(dst, row, col, sum / 4.0f);
}

// LINEAR: reference downsampleX2_linear (matrix), resampling.tpp:422-541. The reference spells the corners,
// border rows, border columns and interior as separate expressions whose tap order differs; the tables hold
// (d_row, d_col) offsets from the anchor in that order. Anchors: corner element / (border row, 2*tc) /
// (2*tr, border column) / (2*tr, 2*tc); far borders mirror the offsets.
static __constant__ signed char c_lin2d_corner[16][2] = { { 0, 0 }, { 1, 0 }, { 0, 1 }, { 1, 1 },
		{ 0, 0 }, { 0, 0 }, { 0, 1 }, { 1, 0 }, { 0, 2 }, { 1, 2 }, { 2, 1 }, { 2, 0 },
		{ 0, 0 }, { 0, 2 }, { 2, 0 }, { 2, 2 } };
static __constant__ signed char c_lin2d_border_row[16][2] = { { 0, 0 }, { 0, 1 }, { 1, 0 }, { 1, 1 },
		{ 0, -1 }, { 0, 0 }, { 0, 1 }, { 0, 2 }, { 1, -1 }, { 2, 0 }, { 2, 1 }, { 1, 2 },
		{ 0, -1 }, { 0, 2 }, { 2, -1 }, { 2, 2 } };
static __constant__ signed char c_lin2d_border_col[16][2] = { { 0, 0 }, { 1, 0 }, { 0, 1 }, { 1, 1 },
		{ -1, 0 }, { 0, 0 }, { 1, 0 }, { 2, 0 }, { -1, 1 }, { 0, 2 }, { 1, 2 }, { 2, 1 },
		{ -1, 0 }, { 2, 0 }, { -1, 2 }, { 2, 2 } };
static __constant__ signed char c_lin2d_interior[16][2] = { { 0, 0 }, { 0, 1 }, { 1, 0 }, { 1, 1 },
		{ -1, 0 }, { 0, -1 }, { -1, 1 }, { 0, 2 }, { 2, 0 }, { 1, -1 }, { 2, 1 }, { 1, 2 },
		{ -1, -1 }, { -1, 2 }, { 2, -1 }, { 2, 2 } };

template<typename Access>
static __global__ void k_downsample_linear2d(Access acc, Grid2 src, Grid2 dst) {
	LSF_PIXEL_2D(dst);
	(void) idx;
	if (!in_grid) return;
	const float coeff0 = 0.140625f, coeff1 = 0.046875f, coeff2 = 0.015625f;
	const bool row_near = row == 0, row_far = row == dst.H - 1;
	const bool col_near = col == 0, col_far = col == dst.W - 1;
	const signed char (*table)[2];
	int r0, c0, sr = 1, sc = 1;
	if ((row_near || row_far) && (col_near || col_far)) {
		table = c_lin2d_corner;
		r0 = row_near ? 0 : src.H - 1;
		c0 = col_near ? 0 : src.W - 1;
		sr = row_near ? 1 : -1;
		sc = col_near ? 1 : -1;
	} else if (row_near || row_far) {
		table = c_lin2d_border_row;
		r0 = row_near ? 0 : src.H - 1;
		sr = row_near ? 1 : -1;
		c0 = 2 * col;
	} else if (col_near || col_far) {
		table = c_lin2d_border_col;
		c0 = col_near ? 0 : src.W - 1;
		sc = col_near ? 1 : -1;
		r0 = 2 * row;
	} else {
		table = c_lin2d_interior;
		r0 = 2 * row;
		c0 = 2 * col;
	}
	auto tap = [&](int t) {return acc.load(src, r0 + sr * table[t][0], c0 + sc * table[t][1]);};
	auto s0 = tap(0);
	for (int t = 1; t < 4; t++) s0 = s0 + tap(t);
	auto s1 = tap(4);
	for (int t = 5; t < 12; t++) s1 = s1 + tap(t);
	auto s2 = tap(12);
	for (int t = 13; t < 16; t++) s2 = s2 + tap(t);
	acc.store(dst, row, col, (coeff0 * s0 + coeff1 * s1) + coeff2 * s2);
}

// ---------------------------------------------------------------------------------------------- prolong x2
// reference upsampleX2_nearest (matrix) resampling.tpp:68-82 / upsampleX2_linear (matrix) :130-216
static __global__ void k_upsample2d(const float* __restrict__ src, float* __restrict__ dst, int channels, Grid2 sg, Grid2 dg,
		int linear) {
	LSF_PIXEL_2D(dg);
	if (!in_grid) return;
	for (int c = 0; c < channels; c++) {
		const float* s = src + c * sg.N;
		dst[c * dg.N + idx] = linear ? upsample_linear2d_at(s, sg.W, 1, sg.H, sg.W, row, col)
				: s[(long long) (row >> 1) * sg.W + (col >> 1)];
	}
}

#endif  // __CUDACC__

}  // namespace lsf
