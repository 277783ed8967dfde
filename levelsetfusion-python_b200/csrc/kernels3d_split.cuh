// kernels3d_split.cuh -- second-generation split of the 3D hierarchical iteration (sm_100a).
//
// ncu of that kernel (capture not kept; figures quoted in profiles/r1_ncu_tma_v3.md) showed the lane-contiguous stage 1 to be ISSUE bound (70 % issue-active, 435
// executed instructions per voxel, a third of them integer / branch overhead) and the single-kernel three-pass filter
// to be latency bound (one 768-thread block per SM because its axis-0 sliding window costs 48 registers per thread,
// barrier + long-scoreboard stalls). The iteration is therefore cut at a different place:
//
//   k_hier_stage1_xpass<TIKHONOV,R>  stage 1 (gather + data term + Tikhonov term, reference optimizer.tpp:186-200) AND
//                                    the axis-0 pass of the separable filter (convolution.cpp:240-267). A thread owns
//                                    one (y,z) column and marches along x; the filter pass is a chain of 2R+1 partial
//                                    sums per component held in registers (each new plane is added to every pending
//                                    output in tap order, which is exactly the reference's accumulation order), so
//                                    it needs no halo, no shared memory and no barrier. Borders are handled with
//                                    clamped loads + selects instead of branches.
//   k_sobolev_yz3d<R>                axis-1 and axis-2 passes on shared-memory tiles + warp update + max-norm
//                                    (convolution.cpp:268-331, optimizer.tpp:207-211); without the register window
//                                    two blocks fit an SM, and the next plane is prefetched while this one is filtered.
//
// HBM traffic is unchanged (56 + 48 B/voxel); arithmetic is unchanged (float32, reference order, no FMA): bit-identical.
#pragma once

#include "kernels3d_fused.cuh"

namespace lsf {

#ifdef __CUDACC__

struct XPassArgs {
	float k[7];   // flipped taps: k[q] multiplies in[i - R + q]
	int x_chunk;  // output planes per block along axis 0
	unsigned long long one2;  // {1.0f, 1.0f}: opaque multiplier of the packed adds (kernels3d_tma.cuh)
	unsigned long long k2[7]; // the taps duplicated into both lanes of an f32x2 (k_hier_stage1_tma's packed chain)
	// k_hier_stage1_tma, one volume: chunks of unequal length (chunk c = planes [bounds[c], bounds[c + 1])), longest first,
	// so that the blocks of the last wave are short ones (marching_schedule); chunk_count == 0: chunks of x_chunk planes
	int chunk_count;
	short chunk_bounds[13];
};

// one axis of the replicated-border Laplacian without branches (reference gradients.tpp:28-35,114-171):
// kind 0 = first element (next - cur), 1 = last (prev - cur), 2 = interior ((next - 2 cur) + prev), 3 = axis of length 1
__device__ __forceinline__ float laplace_select(float prev, float cur, float next, int kind) {
	const float interior = (next - 2.0f * cur) + prev;
	const float low = next - cur, high = prev - cur;
	float r = kind == 0 ? low : interior;
	r = kind == 1 ? high : r;
	return kind == 3 ? 0.0f : r;
}
__device__ __forceinline__ int border_kind(int i, int n) {
	return n < 2 ? 3 : (i == 0 ? 0 : (i == n - 1 ? 1 : 2));
}

template<bool TIKHONOV, int R>
static __global__ void __launch_bounds__(256) k_hier_stage1_xpass(HierIterArgs a, XPassArgs t) {
	constexpr int K = 2 * R + 1;
	if (a.check_convergence && level_converged(a.max_sq_bits, a.iteration, a.threshold)) return;
	const int X = a.g.X, Y = a.g.Y, Z = a.g.Z;
	const int z = blockIdx.x * 32 + threadIdx.x;
	const int y = blockIdx.y * 8 + threadIdx.y;
	if (z >= Z || y >= Y) return;
	const int YZ = Y * Z;
	const int N = (int) a.g.N;
	const int xs = blockIdx.z * t.x_chunk;
	const int xe = min(X, xs + t.x_chunk);
	float k[K];
#pragma unroll
	for (int q = 0; q < K; q++) k[q] = t.k[q];
	float acc[3][K];
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int q = 0; q < K; q++) acc[c][q] = 0.0f;

	// in-plane neighbour offsets, clamped at the borders (the clamped value is never selected there)
	const int off_ym = y > 0 ? -Z : 0, off_yp = y + 1 < Y ? Z : 0;
	const int off_zm = z > 0 ? -1 : 0, off_zp = z + 1 < Z ? 1 : 0;
	const int kind_y = border_kind(y, Y), kind_z = border_kind(z, Z);

	const int x_first = max(xs - R, 0);  // planes below 0 contribute zeros to accumulators that are still zero
	const int x_last = xe + R - 1;       // planes >= X contribute zeros
	int idx = x_first * YZ + y * Z + z;
	float prev[3] = { 0.f, 0.f, 0.f }, cur[3] = { 0.f, 0.f, 0.f };
	if (TIKHONOV) {
#pragma unroll
		for (int c = 0; c < 3; c++) {
			cur[c] = __ldg(a.g_prev + c * N + idx);
			prev[c] = __ldg(a.g_prev + c * N + idx - (x_first > 0 ? YZ : 0));
		}
	}
#pragma unroll 1
	for (int x = x_first; x <= x_last; x++, idx += YZ) {
		float g[3] = { 0.f, 0.f, 0.f };
		if (x < X) {
			const float wx = __ldg(a.warp + idx), wy = __ldg(a.warp + N + idx), wz = __ldg(a.warp + 2 * N + idx);
			const float cn = __ldg(a.canonical + idx);
			float lap[3] = { 0.f, 0.f, 0.f };
			if (TIKHONOV) {
				const int off_xp = x + 1 < X ? YZ : 0;
				const int kind_x = border_kind(x, X);
#pragma unroll
				for (int c = 0; c < 3; c++) {
					const float* p = a.g_prev + c * N + idx;
					const float next = __ldg(p + off_xp);
					const float ym = __ldg(p + off_ym), yp = __ldg(p + off_yp);
					const float zm = __ldg(p + off_zm), zp = __ldg(p + off_zp);
					float acc_l = laplace_select(prev[c], cur[c], next, kind_x);
					acc_l += laplace_select(ym, cur[c], yp, kind_y);
					acc_l += laplace_select(zm, cur[c], zp, kind_z);
					lap[c] = acc_l;
					prev[c] = cur[c];
					cur[c] = next;
				}
			}
			const float4 s = gather4i(a.pack, X, Y, Z, x, y, z, wx, wy, wz);
			const float diff = s.x - cn;
			g[0] = (s.y * diff) * a.amplifier;
			g[1] = (s.z * diff) * a.amplifier;
			g[2] = (s.w * diff) * a.amplifier;
			if (TIKHONOV) {
				g[0] = g[0] - lap[0] * a.strength;
				g[1] = g[1] - lap[1] * a.strength;
				g[2] = g[2] - lap[2] * a.strength;
			}
		}
		// axis-0 filter pass: plane x is tap q of output plane x + R - q; acc[c][q] holds the partial sum (taps 0..q) of
		// output x + R - q, so adding in place from the oldest output down reproduces sum_{q ascending} in[.]*k[q]
#pragma unroll
		for (int c = 0; c < 3; c++) {
#pragma unroll
			for (int q = K - 1; q >= 1; q--) acc[c][q] = acc[c][q - 1] + g[c] * k[q];
			acc[c][0] = g[c] * k[0];
		}
		const int xo = x - R;
		if (xo >= xs) {
			const int out = idx - R * YZ;
			a.g_out[out] = acc[0][K - 1];
			a.g_out[N + out] = acc[1][K - 1];
			a.g_out[2 * N + out] = acc[2][K - 1];
		}
	}
}

// ---------------------------------------------------------------------------------------------- axis-1 / axis-2 passes
// Input: the field after the axis-0 pass. Same tile geometry and pass code as k_sobolev_fused3d.
template<int R>
static __global__ void __launch_bounds__(FusedConv<R>::THREADS, 2) k_sobolev_yz3d(FusedConvArgs a) {
	typedef FusedConv<R> C;
	constexpr int K = C::K;
	if (a.check_convergence && level_converged(a.max_sq_bits, a.iteration, a.threshold)) return;

	__shared__ float s1[3][C::COLS];           // after the axis-0 pass, extended tile (input of this kernel)
	__shared__ float s2[3][C::TY * C::S2];     // after the axis-1 pass, TY x EZ
	__shared__ float s3[3][C::TY * C::S3];     // after the axis-2 pass, TY x TZ

	const int Y = a.g.Y, Z = a.g.Z;
	const int YZ = Y * Z;
	const int tid = threadIdx.x;
	const int c = tid / C::GROUP;
	const int t = tid - c * C::GROUP;
	const int z0 = blockIdx.x * C::TZ, y0 = blockIdx.y * C::TY;
	const int x0 = a.x_begin + blockIdx.z * a.x_chunk;
	const int x1 = min(a.x_end, x0 + a.x_chunk);
	float k[K];
#pragma unroll
	for (int q = 0; q < K; q++) k[q] = a.k[q];

	const float* plane = a.in + (long long) c * a.g.N + (long long) x0 * YZ;
	int column[C::CPT];
	unsigned inside = 0;
#pragma unroll
	for (int j = 0; j < C::CPT; j++) {
		const int col = t + j * C::GROUP;
		const int yy = col / C::EZ, zz = col - yy * C::EZ;
		const int gy = y0 - R + yy, gz = z0 - R + zz;
		const bool ok = col < C::COLS && gy >= 0 && gy < Y && gz >= 0 && gz < Z;
		inside |= ok ? (1u << j) : 0u;
		column[j] = ok ? gy * Z + gz : 0;
	}
	float next[C::CPT];  // the plane being prefetched
#pragma unroll
	for (int j = 0; j < C::CPT; j++) next[j] = ((inside >> j) & 1u) ? __ldg(plane + column[j]) : 0.0f;

	const int p2_zz = t % C::EZ, p2_yb = t / C::EZ;
	const float* s1_read = &s1[c][p2_yb * C::YB * C::EZ + p2_zz];
	float* s2_write = &s2[c][p2_yb * C::YB * C::S2 + p2_zz];
	const int p3_y = t & (C::TY - 1), p3_zb = t / C::TY;
	const float* s2_read = &s2[c][p3_y * C::S2 + p3_zb * C::ZB];
	float* s3_write = &s3[c][p3_y * C::S3 + p3_zb * C::ZB];
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

	for (int x = x0; x < x1; x++) {
		// s1 of the previous plane was last read before the second barrier of the previous iteration
#pragma unroll
		for (int j = 0; j < C::CPT; j++)
			if (j < C::CPT - 1 || t + j * C::GROUP < C::COLS) s1[c][t + j * C::GROUP] = next[j];
		__syncthreads();
		plane += YZ;
		if (x + 1 < x1) {
#pragma unroll
			for (int j = 0; j < C::CPT; j++) next[j] = ((inside >> j) & 1u) ? __ldg(plane + column[j]) : 0.0f;
		}
		// ---- axis-1 pass
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
		// ---- axis-2 pass
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
		// ---- write back: filtered gradient, warp update, max ||g||^2
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
		// s3 is next written after two more barriers; s1 after the next loop head, s2 after the next first barrier
	}
	if (a.max_sq_bits != nullptr) block_atomic_max(best, a.max_sq_bits + a.iteration);
}

// ---------------------------------------------------------------------------------------------- host-side launch helpers
template<int R>
void launch_split_iteration(bool tikhonov, HierIterArgs a, const Taps& taps, float* h, float* filtered, float* warp,
		int x_chunk_stage1, int x_chunk_filter, cudaStream_t stream, cudaEvent_t* events) {
	typedef FusedConv<R> C;
	const Grid3& g = a.g;
	XPassArgs t;
	for (int q = 0; q < 7; q++) t.k[q] = q < C::K ? taps.k[q] : 0.0f;
	t.x_chunk = x_chunk_stage1;
	t.one2 = 0x3f8000003f800000ull;
	a.g_out = h;
	const dim3 block(32, 8, 1), grid(div_up(g.Z, 32), div_up(g.Y, 8), div_up(g.X, t.x_chunk));
	if (tikhonov) k_hier_stage1_xpass<true, R> <<<counted(grid), block, 0, stream>>>(a, t);
	else k_hier_stage1_xpass<false, R> <<<counted(grid), block, 0, stream>>>(a, t);
	if (events) cudaEventRecord(events[1], stream);
	FusedConvArgs f;
	f.in = h;
	f.out = filtered;
	f.warp = warp;
	f.g = g;
	for (int q = 0; q < 7; q++) f.k[q] = q < C::K ? taps.k[q] : 0.0f;
	f.rate = a.rate;
	f.threshold = a.threshold;
	f.max_sq_bits = a.max_sq_bits;
	f.iteration = a.iteration;
	f.check_convergence = a.check_convergence;
	f.x_begin = 0;
	f.x_end = g.X;
	f.x_chunk = x_chunk_filter;
	const dim3 fgrid(div_up(g.Z, C::TZ), div_up(g.Y, C::TY), div_up(g.X, f.x_chunk));
	k_sobolev_yz3d<R> <<<counted(fgrid), C::THREADS, 0, stream>>>(f);
	if (events) cudaEventRecord(events[2], stream);
}

#endif  // __CUDACC__

}  // namespace lsf
