// kernels3d_fused.cuh -- the two production kernels of one 3D hierarchical iteration (sm_100a).
//
// The ncu capture of the first (unfused) version (profiles/r1_ncu_full_baseline_unfused.md) showed every stage to
// be issue-bound (440-600 executed instructions per voxel) with DRAM at 11-29 % of peak. These kernels cut the
// instruction count per voxel and the number of passes over HBM:
//
//   k_hier_gradient3d_v4   stage 1 (gather + data term + Tikhonov term [+ update + max-norm]) on 4 consecutive z
//                          voxels per thread: 128-bit loads/stores of warp, canonical, g; 32-bit index arithmetic.
//   k_sobolev_fused3d<R>   stage 2: ALL THREE passes of the separable filter + warp update + max-norm in one kernel.
//                          A block owns a 32x32 (y,z) tile and marches along x. The axis-0 pass keeps a
//                          (2R+1)-deep sliding window per (y,z) column in REGISTERS (one coalesced LDG per column
//                          and plane, software-prefetched one plane ahead); the axis-1 and axis-2 passes run on
//                          shared-memory planes with register blocking (8 resp. 4 outputs per thread). The three
//                          vector components are handled by three groups of 256 threads so the max-norm
//                          (needs all components of a voxel) is formed in-block. HBM traffic: read g_pre once
//                          (+ tile halo, served by L2), read/write warp once, write g once = 48 B/voxel instead
//                          of 96 B/voxel for three separate passes.
//
// Arithmetic is unchanged (float32, reference operation order, no FMA): results stay bit-identical to the oracle.
#pragma once

#include "kernels3d.cuh"

#include <algorithm>

namespace lsf {

#ifdef __CUDACC__

__device__ __forceinline__ float4 ld4(const float* p) {
	return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st4(float* p, float4 v) {
	*reinterpret_cast<float4*>(p) = v;
}
__device__ __forceinline__ float4 operator-(float4 a, float4 b) {
	return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
}

// trilinear gather with 32-bit index arithmetic (see gather4 in kernels3d.cuh for the reference citations)
__device__ __forceinline__ float4 gather4i(const float4* __restrict__ pack, int X, int Y, int Z, int x, int y, int z,
		float wx, float wy, float wz) {
	const float lookup_x = (float) x + wx;
	const float lookup_y = (float) y + wy;
	const float lookup_z = (float) z + wz;
	int bx = __float2int_rd(lookup_x);
	int by = __float2int_rd(lookup_y);
	int bz = __float2int_rd(lookup_z);
	const float rx = lookup_x - (float) bx, ry = lookup_y - (float) by, rz = lookup_z - (float) bz;
	const float ix = 1.0f - rx, iy = 1.0f - ry, iz = 1.0f - rz;
	bx = min(max(bx, -2), X);
	by = min(max(by, -2), Y);
	bz = min(max(bz, -2), Z);
	const int sy = Z + 4, sx = (Y + 4) * (Z + 4);
	const float4* p = pack + ((bx + 2) * sx + (by + 2) * sy + (bz + 2));
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

// The same gather for a rank that holds only planes [pack_origin, pack_origin + pack_X) of the level's pack (slab
// decomposition). The lookup is formed from the GLOBAL plane index so that the interpolation ratios are bit-identical
// to the whole-volume run; a tap that would fall into another rank's part of the pack raises `violation`.
__device__ __forceinline__ float4 gather4i_slab(const float4* __restrict__ pack, const HierIterArgs& a, int x_global,
		int y, int z, float wx, float wy, float wz) {
	const int Y = a.g.Y, Z = a.g.Z;
	const float lookup_x = (float) x_global + wx;
	const float lookup_y = (float) y + wy;
	const float lookup_z = (float) z + wz;
	int bx = __float2int_rd(lookup_x);
	int by = __float2int_rd(lookup_y);
	int bz = __float2int_rd(lookup_z);
	const float rx = lookup_x - (float) bx, ry = lookup_y - (float) by, rz = lookup_z - (float) bz;
	const float ix = 1.0f - rx, iy = 1.0f - ry, iz = 1.0f - rz;
	bx = min(max(bx, -2), a.X_global) - a.pack_origin;
	if ((a.pack_interior_low && bx < 0) || (a.pack_interior_high && bx + 1 > a.pack_X - 1)) {
		if (a.violation != nullptr) *a.violation = 1;
	}
	bx = min(max(bx, -2), a.pack_X);
	by = min(max(by, -2), Y);
	bz = min(max(bz, -2), Z);
	const int sy = Z + 4, sx = (Y + 4) * (Z + 4);
	const float4* p = pack + ((bx + 2) * sx + (by + 2) * sy + (bz + 2));
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

// replicated-border Laplacian of one plane at 4 consecutive z voxels (reference gradients.tpp:28-35,114-171)
__device__ __forceinline__ void laplacian4(const float* __restrict__ p, int idx, int x, int y, int z, int X, int Y,
		int Z, float (&out)[4]) {
	const int YZ = Y * Z;
	const float4 c = ld4(p + idx);
	float4 tx = make_float4(0.f, 0.f, 0.f, 0.f), ty = tx;
	if (X >= 2) {
		if (x == 0) tx = ld4(p + idx + YZ) - c;
		else if (x == X - 1) tx = ld4(p + idx - YZ) - c;
		else tx = (ld4(p + idx + YZ) - 2.0f * c) + ld4(p + idx - YZ);
	}
	if (Y >= 2) {
		if (y == 0) ty = ld4(p + idx + Z) - c;
		else if (y == Y - 1) ty = ld4(p + idx - Z) - c;
		else ty = (ld4(p + idx + Z) - 2.0f * c) + ld4(p + idx - Z);
	}
	const float left = z > 0 ? __ldg(p + idx - 1) : 0.0f;
	const float right = z + 4 < Z ? __ldg(p + idx + 4) : 0.0f;
	const float cv[6] = { left, c.x, c.y, c.z, c.w, right };
	const float txv[4] = { tx.x, tx.y, tx.z, tx.w }, tyv[4] = { ty.x, ty.y, ty.z, ty.w };
#pragma unroll
	for (int v = 0; v < 4; v++) {
		const int i = z + v;
		float tz;
		if (i == 0) tz = cv[v + 2] - cv[v + 1];
		else if (i == Z - 1) tz = cv[v] - cv[v + 1];
		else tz = (cv[v + 2] - 2.0f * cv[v + 1]) + cv[v];
		out[v] = (txv[v] + tyv[v]) + tz;
	}
}

// ---------------------------------------------------------------------------------------------- stage 1, 4 voxels/thread
// Same arithmetic as k_hier_gradient3d (reference optimizer.tpp:186-200 [+ :207-211]). Requires Z % 4 == 0 (>= 4),
// 16-byte aligned planes and voxel counts below 2^31 / 3.
template<bool TIKHONOV, bool FUSE_UPDATE>
static __global__ void __launch_bounds__(256) k_hier_gradient3d_v4(HierIterArgs a) {
	if (a.check_convergence && level_converged(a.max_sq_bits, a.iteration, a.threshold)) return;
	const int X = a.g.X, Y = a.g.Y, Z = a.g.Z;
	const int z = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
	const int y = blockIdx.y * blockDim.y + threadIdx.y;
	const int x = blockIdx.z * blockDim.z + threadIdx.z;
	float best = 0.0f;
	if (z < Z && y < Y && x < X) {
		const int idx = (x * Y + y) * Z + z;
		const long long N = a.g.N;
		const float4 wx4 = ld4(a.warp + idx), wy4 = ld4(a.warp + N + idx), wz4 = ld4(a.warp + 2 * N + idx);
		const float4 cn4 = ld4(a.canonical + idx);
		// the Tikhonov loads do not depend on the warp: issue them first so they overlap the dependent gather chain
		float lap_x[4], lap_y[4], lap_z[4];
		if (TIKHONOV) {
			laplacian4(a.g_prev, idx, x, y, z, X, Y, Z, lap_x);
			laplacian4(a.g_prev + N, idx, x, y, z, X, Y, Z, lap_y);
			laplacian4(a.g_prev + 2 * N, idx, x, y, z, X, Y, Z, lap_z);
		}
		const float wx[4] = { wx4.x, wx4.y, wx4.z, wx4.w }, wy[4] = { wy4.x, wy4.y, wy4.z, wy4.w };
		const float wz[4] = { wz4.x, wz4.y, wz4.z, wz4.w }, cn[4] = { cn4.x, cn4.y, cn4.z, cn4.w };
		float gx[4], gy[4], gz[4];
#pragma unroll
		for (int v = 0; v < 4; v++) {
			const float4 s = gather4i(a.pack, X, Y, Z, x, y, z + v, wx[v], wy[v], wz[v]);
			const float diff = s.x - cn[v];
			gx[v] = (s.y * diff) * a.amplifier;
			gy[v] = (s.z * diff) * a.amplifier;
			gz[v] = (s.w * diff) * a.amplifier;
			if (TIKHONOV) {
				gx[v] = gx[v] - lap_x[v] * a.strength;
				gy[v] = gy[v] - lap_y[v] * a.strength;
				gz[v] = gz[v] - lap_z[v] * a.strength;
			}
		}
		if (a.g_out != nullptr) {
			st4(a.g_out + idx, make_float4(gx[0], gx[1], gx[2], gx[3]));
			st4(a.g_out + N + idx, make_float4(gy[0], gy[1], gy[2], gy[3]));
			st4(a.g_out + 2 * N + idx, make_float4(gz[0], gz[1], gz[2], gz[3]));
		}
		if (FUSE_UPDATE) {
			st4(a.warp_out + idx, make_float4(wx[0] - gx[0] * a.rate, wx[1] - gx[1] * a.rate, wx[2] - gx[2] * a.rate,
					wx[3] - gx[3] * a.rate));
			st4(a.warp_out + N + idx, make_float4(wy[0] - gy[0] * a.rate, wy[1] - gy[1] * a.rate,
					wy[2] - gy[2] * a.rate, wy[3] - gy[3] * a.rate));
			st4(a.warp_out + 2 * N + idx, make_float4(wz[0] - gz[0] * a.rate, wz[1] - gz[1] * a.rate,
					wz[2] - gz[2] * a.rate, wz[3] - gz[3] * a.rate));
#pragma unroll
			for (int v = 0; v < 4; v++) {
				float sq = gx[v] * gx[v];
				sq += gy[v] * gy[v];
				sq += gz[v] * gz[v];
				if (sq > best) best = sq;
			}
		}
	}
	if (FUSE_UPDATE) block_atomic_max(best, a.max_sq_bits + a.iteration);
}

inline void launch_shape_v4(const Grid3& g, dim3* grid, dim3* block) {
	const int zgroups = g.Z / 4;
	int bx = 1, by = 1, bz = 1;
	while (bx < zgroups && bx < 32) bx *= 2;
	while (by < g.Y && bx * by < 256) by *= 2;
	while (bz < g.X && bx * by * bz < 256) bz *= 2;
	*block = dim3(bx, by, bz);
	*grid = dim3(div_up(zgroups, bx), div_up(g.Y, by), div_up(g.X, bz));
}

// ---------------------------------------------------------------------------------------------- stage 1, lane-contiguous
// Same arithmetic again (reference optimizer.tpp:186-200 [+ :207-211]), organised for the L1 pipe: the ncu capture of
// the 4-voxel version (profiles/r1_ncu_fused_v1.md) showed 25 sectors per gather request because the lanes of a warp
// addressed pack entries 64 B apart. Here the 32 lanes of a warp are 32 consecutive z voxels (every LDG.128 of the
// gather touches 4 lines, every scalar load one), and a thread marches over XV planes along x so that the Tikhonov
// term re-uses the x-1 / x / x+1 values of g_prev from registers.
__device__ __forceinline__ float laplace_axis3(float prev, float cur, float next, int i, int n) {
	if (n < 2) return 0.0f;
	if (i == 0) return next - cur;
	if (i == n - 1) return prev - cur;
	return (next - 2.0f * cur) + prev;
}

template<bool TIKHONOV, bool FUSE_UPDATE, int XV>
static __global__ void __launch_bounds__(256) k_hier_gradient3d_lane(HierIterArgs a) {
	if (a.check_convergence && level_converged(a.max_sq_bits, a.iteration, a.threshold)) return;
	const int X = a.g.X, Y = a.g.Y, Z = a.g.Z;
	const int YZ = Y * Z;
	const int N = (int) a.g.N;
	const int z = blockIdx.x * 32 + threadIdx.x;
	const int y = blockIdx.y * 8 + threadIdx.y;
	const int xb = a.x_begin + blockIdx.z * XV;
	float best = 0.0f;
	if (z < Z && y < Y) {
		int idx = (xb * Y + y) * Z + z;
		float prev[3] = { 0.f, 0.f, 0.f }, cur[3] = { 0.f, 0.f, 0.f };
		if (TIKHONOV) {
#pragma unroll
			for (int c = 0; c < 3; c++) {
				cur[c] = __ldg(a.g_prev + c * N + idx);
				if (xb > 0) prev[c] = __ldg(a.g_prev + c * N + idx - YZ);
			}
		}
#pragma unroll
		for (int v = 0; v < XV; v++) {
			const int x = xb + v;
			if (x >= a.x_end) break;
			const float wx = __ldg(a.warp + idx), wy = __ldg(a.warp + N + idx), wz = __ldg(a.warp + 2 * N + idx);
			const float cn = __ldg(a.canonical + idx);
			float lap[3] = { 0.f, 0.f, 0.f };
			if (TIKHONOV) {
#pragma unroll
				for (int c = 0; c < 3; c++) {
					const float* p = a.g_prev + c * N + idx;
					const float next = (x + 1 < X) ? __ldg(p + YZ) : 0.0f;
					const float yp = (y > 0) ? __ldg(p - Z) : 0.0f, yn = (y + 1 < Y) ? __ldg(p + Z) : 0.0f;
					const float zp = (z > 0) ? __ldg(p - 1) : 0.0f, zn = (z + 1 < Z) ? __ldg(p + 1) : 0.0f;
					float acc = laplace_axis3(prev[c], cur[c], next, x, X);
					acc += laplace_axis3(yp, cur[c], yn, y, Y);
					acc += laplace_axis3(zp, cur[c], zn, z, Z);
					lap[c] = acc;
					prev[c] = cur[c];
					cur[c] = next;
				}
			}
			const float4 s = gather4i_slab(a.pack, a, x + a.x_origin, y, z, wx, wy, wz);
			const float diff = s.x - cn;
			float gx = (s.y * diff) * a.amplifier;
			float gy = (s.z * diff) * a.amplifier;
			float gz = (s.w * diff) * a.amplifier;
			if (TIKHONOV) {
				gx = gx - lap[0] * a.strength;
				gy = gy - lap[1] * a.strength;
				gz = gz - lap[2] * a.strength;
			}
			if (a.g_out != nullptr) {
				a.g_out[idx] = gx;
				a.g_out[N + idx] = gy;
				a.g_out[2 * N + idx] = gz;
			}
			if (FUSE_UPDATE) {
				a.warp_out[idx] = wx - gx * a.rate;
				a.warp_out[N + idx] = wy - gy * a.rate;
				a.warp_out[2 * N + idx] = wz - gz * a.rate;
				float sq = gx * gx;
				sq += gy * gy;
				sq += gz * gz;
				if (sq > best) best = sq;
			}
			idx += YZ;
		}
	}
	if (FUSE_UPDATE) block_atomic_max(best, a.max_sq_bits + a.iteration);
}

inline void launch_shape_lane(const Grid3& g, int x_planes, int xv, dim3* grid, dim3* block) {
	*block = dim3(32, 8, 1);
	*grid = dim3(div_up(g.Z, 32), div_up(g.Y, 8), div_up(x_planes, xv));
}

// ---------------------------------------------------------------------------------------------- stage 2, fused
template<int R>
struct FusedConv {
	static constexpr int K = 2 * R + 1;
	static constexpr int TY = 32, TZ = 32;            // output tile (y, z)
	static constexpr int EY = TY + 2 * R, EZ = TZ + 2 * R;  // tile + filter halo
	static constexpr int GROUP = 256;                 // threads per vector component
	static constexpr int THREADS = 3 * GROUP;
	static constexpr int COLS = EY * EZ;              // axis-0 pass columns per component
	static constexpr int CPT = (COLS + GROUP - 1) / GROUP;  // columns per thread
	static constexpr int YB = 8;                      // axis-1 pass: outputs per task
	static constexpr int P2_TASKS = EZ * (TY / YB);
	static constexpr int ZB = 4;                      // axis-2 pass: outputs per task
	static constexpr int S2 = EZ | 1;                 // odd row stride: lanes along y are conflict-free
	static constexpr int S3 = TZ + 1;
	static_assert(P2_TASKS <= GROUP, "axis-1 pass must fit one round");
	static_assert(TY * (TZ / ZB) == GROUP, "axis-2 pass is exactly one round");
};

struct FusedConvArgs {
	const float* in;   // planes: gradient before filtering
	float* out;        // planes: filtered gradient (may be nullptr when nothing reads it)
	float* warp;       // planes: updated in place, warp -= out * rate (nullptr: filter only)
	Grid3 g;
	float k[7];        // flipped taps: k[q] multiplies in[i - R + q]
	float rate, threshold;
	unsigned* max_sq_bits;
	int iteration;
	int check_convergence;
	int x_chunk;       // planes per block along axis 0
	int x_begin, x_end;  // planes filtered / updated (whole volume: 0, X); loads still see every plane of the allocation
};

// reference convolve_with_kernel (tensor), cpp/src/math/convolution.cpp:221-332, followed by
// optimizer.tpp:207-211 (warp update, max-norm); zero padding = out-of-volume loads return 0.
template<int R>
static __global__ void __launch_bounds__(FusedConv<R>::THREADS, 1) k_sobolev_fused3d(FusedConvArgs a) {
	typedef FusedConv<R> C;
	constexpr int K = C::K;
	constexpr int RING = K + 1;  // window slots per column: K live planes + the one being prefetched
	if (a.check_convergence && level_converged(a.max_sq_bits, a.iteration, a.threshold)) return;

	__shared__ float s1[3][C::COLS];           // after the axis-0 pass, extended tile
	__shared__ float s2[3][C::TY * C::S2];     // after the axis-1 pass, TY x EZ
	__shared__ float s3[3][C::TY * C::S3];     // after the axis-2 pass, TY x TZ

	const int X = a.g.X, Y = a.g.Y, Z = a.g.Z;
	const int YZ = Y * Z;
	const int tid = threadIdx.x;
	const int c = tid / C::GROUP;  // vector component of this thread (warp-uniform)
	const int t = tid - c * C::GROUP;
	const int z0 = blockIdx.x * C::TZ, y0 = blockIdx.y * C::TY;
	const int x0 = a.x_begin + blockIdx.z * a.x_chunk;
	const int x1 = min(a.x_end, x0 + a.x_chunk);
	float k[K];
#pragma unroll
	for (int q = 0; q < K; q++) k[q] = a.k[q];

	// axis-0 pass bookkeeping: one plane pointer (advanced by one plane per step) + per-column offsets
	const float* plane = a.in + (long long) c * a.g.N + (long long) (x0 - R) * YZ;
	int column[C::CPT];
	unsigned inside = 0;  // bit j: column j lies inside the volume in y and z
#pragma unroll
	for (int j = 0; j < C::CPT; j++) {
		const int col = t + j * C::GROUP;
		const int yy = col / C::EZ, zz = col - yy * C::EZ;
		const int gy = y0 - R + yy, gz = z0 - R + zz;
		const bool ok = col < C::COLS && gy >= 0 && gy < Y && gz >= 0 && gz < Z;
		inside |= ok ? (1u << j) : 0u;
		column[j] = ok ? gy * Z + gz : 0;
	}
	float w[C::CPT][RING];  // sliding windows (registers); slot (s % RING) holds plane x0 - R + s
	// planes x0-R .. x0+R-1 -> slots 0 .. K-2; plane x0+R -> slot K-1 (the "prefetched" one of step 0)
#pragma unroll
	for (int p = 0; p < K; p++) {
		const bool plane_ok = (x0 - R + p) >= 0 && (x0 - R + p) < X;
#pragma unroll
		for (int j = 0; j < C::CPT; j++)
			w[j][p] = (plane_ok && ((inside >> j) & 1u)) ? __ldg(plane + column[j]) : 0.0f;
		plane += YZ;
	}  // plane now points at plane x0 + R + 1

	// axis-1 / axis-2 pass task coordinates (constant per thread)
	const int p2_zz = t % C::EZ, p2_yb = t / C::EZ;
	const float* s1_read = &s1[c][p2_yb * C::YB * C::EZ + p2_zz];
	float* s2_write = &s2[c][p2_yb * C::YB * C::S2 + p2_zz];
	const int p3_y = t & (C::TY - 1), p3_zb = t / C::TY;
	const float* s2_read = &s2[c][p3_y * C::S2 + p3_zb * C::ZB];
	float* s3_write = &s3[c][p3_y * C::S3 + p3_zb * C::ZB];
	// write-back voxels of this thread: v = tid and tid + THREADS (the second only for tid < TY*TZ - THREADS)
	constexpr int WB = (C::TY * C::TZ + C::THREADS - 1) / C::THREADS;
	int wb_s3[WB], wb_global[WB];
	bool wb_ok[WB];
#pragma unroll
	for (int i = 0; i < WB; i++) {
		const int v = tid + i * C::THREADS;
		const int yy = v / C::TZ, zz = v - yy * C::TZ;
		wb_ok[i] = v < C::TY * C::TZ && (y0 + yy) < Y && (z0 + zz) < Z;
		wb_s3[i] = yy * C::S3 + zz;
		wb_global[i] = (y0 + yy) * Z + (z0 + zz);
	}
	const long long N = a.g.N;
	float best = 0.0f;

	for (int xb = x0; xb < x1; xb += RING) {
#pragma unroll
		for (int phase = 0; phase < RING; phase++) {
			const int x = xb + phase;
			if (x >= x1) break;
			// ---- axis-0 pass. Live planes x-R..x+R sit in slots phase .. phase+K-1 (mod RING); prefetch plane
			//      x+R+1 into the free slot (phase + K) % RING for the next step.
			{
				const bool plane_ok = (x + R + 1) < X && (x + 1) < x1;
#pragma unroll
				for (int j = 0; j < C::CPT; j++) {
					w[j][(phase + K) % RING] = (plane_ok && ((inside >> j) & 1u)) ? __ldg(plane + column[j]) : 0.0f;
				}
				plane += YZ;
			}
#pragma unroll
			for (int j = 0; j < C::CPT; j++) {
				float acc = w[j][phase % RING] * k[0];
#pragma unroll
				for (int q = 1; q < K; q++) acc += w[j][(phase + q) % RING] * k[q];
				if (j < C::CPT - 1 || t + j * C::GROUP < C::COLS) s1[c][t + j * C::GROUP] = acc;
			}
			__syncthreads();
			// ---- axis-1 pass: YB outputs per task from YB + 2R rows
			if (t < C::P2_TASKS) {
				float v[C::YB + 2 * R];
#pragma unroll
				for (int i = 0; i < C::YB + 2 * R; i++) v[i] = s1_read[i * C::EZ];
#pragma unroll
				for (int i = 0; i < C::YB; i++) {
					float acc = v[i] * k[0];
#pragma unroll
					for (int q = 1; q < K; q++) acc += v[i + q] * k[q];
					s2_write[i * C::S2] = acc;
				}
			}
			__syncthreads();
			// ---- axis-2 pass: ZB outputs per task from ZB + 2R columns (lanes run along y)
			{
				float v[C::ZB + 2 * R];
#pragma unroll
				for (int i = 0; i < C::ZB + 2 * R; i++) v[i] = s2_read[i];
#pragma unroll
				for (int i = 0; i < C::ZB; i++) {
					float acc = v[i] * k[0];
#pragma unroll
					for (int q = 1; q < K; q++) acc += v[i + q] * k[q];
					s3_write[i] = acc;
				}
			}
			__syncthreads();
			// ---- write back: filtered gradient, warp update, max ||g||^2 (all three components per voxel)
#pragma unroll
			for (int i = 0; i < WB; i++) {
				if (wb_ok[i]) {
					const float g0 = s3[0][wb_s3[i]], g1 = s3[1][wb_s3[i]], g2 = s3[2][wb_s3[i]];
					const long long idx = (long long) x * YZ + wb_global[i];
					if (a.out != nullptr) {
						a.out[idx] = g0;
						a.out[N + idx] = g1;
						a.out[2 * N + idx] = g2;
					}
					if (a.warp != nullptr) {
						const float w0 = a.warp[idx], w1 = a.warp[N + idx], w2 = a.warp[2 * N + idx];
						a.warp[idx] = w0 - g0 * a.rate;
						a.warp[N + idx] = w1 - g1 * a.rate;
						a.warp[2 * N + idx] = w2 - g2 * a.rate;
					}
					float sq = g0 * g0;
					sq += g1 * g1;
					sq += g2 * g2;
					if (sq > best) best = sq;
				}
			}
			// no barrier needed here: s3 is next written after two more barriers, s1 after all threads passed
			// the axis-1 barrier of this plane
		}
	}
	if (a.max_sq_bits != nullptr) block_atomic_max(best, a.max_sq_bits + a.iteration);
}

// ---------------------------------------------------------------------------------------------- host-side launch helpers
// Chooses how many planes along axis 0 each block of the fused filter kernel marches over: enough blocks to fill
// whole waves of the GPU, as few as possible because every chunk re-reads 2R priming planes.
template<int R>
int choose_x_chunk(const Grid3& g, int x_planes) {
	typedef FusedConv<R> C;
	static int slots = 0;
	if (slots == 0) {
		int device = 0, sms = 148, per_sm = 1;
		cudaGetDevice(&device);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sobolev_fused3d<R>, C::THREADS, 0);
		slots = sms * std::max(per_sm, 1);
	}
	const long long tiles = (long long) div_up(g.Y, C::TY) * div_up(g.Z, C::TZ);
	int best_chunk = x_planes;
	double best_cost = 1e30;
	for (int chunks = 1; chunks <= std::max(1, x_planes / 8); chunks++) {
		const int chunk = (x_planes + chunks - 1) / chunks;
		const long long blocks = tiles * ((x_planes + chunk - 1) / chunk);
		const double waves = (double) ((blocks + slots - 1) / slots);
		const double cost = waves * (chunk + 2 * R * 0.35);
		if (cost < best_cost - 1e-9) {
			best_cost = cost;
			best_chunk = chunk;
		}
	}
	return best_chunk;
}

template<int R>
void launch_fused_filter(const Taps& taps, float rate, float threshold, const Grid3& g, const float* in, float* out,
		float* warp, unsigned* max_sq_bits, int iteration, int check, cudaStream_t stream, int x_begin, int x_end) {
	typedef FusedConv<R> C;
	FusedConvArgs f;
	f.in = in;
	f.out = out;
	f.warp = warp;
	f.g = g;
	for (int q = 0; q < 7; q++) f.k[q] = q < C::K ? taps.k[q] : 0.0f;
	f.rate = rate;
	f.threshold = threshold;
	f.max_sq_bits = max_sq_bits;
	f.iteration = iteration;
	f.check_convergence = check;
	f.x_begin = x_begin;
	f.x_end = x_end < 0 ? g.X : x_end;
	f.x_chunk = choose_x_chunk<R>(g, f.x_end - f.x_begin);
	const dim3 grid(div_up(g.Z, C::TZ), div_up(g.Y, C::TY), div_up(f.x_end - f.x_begin, f.x_chunk));
	k_sobolev_fused3d<R> <<<counted(grid), C::THREADS, 0, stream>>>(f);
}

// dispatch on the filter radius (1, 2 or 3); returns false if the radius has no fused instantiation
inline bool launch_fused_filter_any(const Taps& taps, float rate, float threshold, const Grid3& g, const float* in,
		float* out, float* warp, unsigned* max_sq_bits, int iteration, int check, cudaStream_t stream, int x_begin = 0,
		int x_end = -1) {
	switch (taps.radius) {
	case 1:
		launch_fused_filter<1>(taps, rate, threshold, g, in, out, warp, max_sq_bits, iteration, check, stream, x_begin, x_end);
		return true;
	case 2:
		launch_fused_filter<2>(taps, rate, threshold, g, in, out, warp, max_sq_bits, iteration, check, stream, x_begin, x_end);
		return true;
	case 3:
		launch_fused_filter<3>(taps, rate, threshold, g, in, out, warp, max_sq_bits, iteration, check, stream, x_begin, x_end);
		return true;
	default:
		return false;
	}
}

#endif  // __CUDACC__

}  // namespace lsf
