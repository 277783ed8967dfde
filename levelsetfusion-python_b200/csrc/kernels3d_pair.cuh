// kernels3d_pair.cuh -- fourth generation of stage 1 of the 3D hierarchical iteration (sm_100a): two z voxels per thread.
//
// ncu of the third generation (profiles/r1_ncu_tma_v3.md): 328 executed thread-instructions per voxel, issue slots 56 %
// busy, DRAM at 40 % -- and an L1 prefetch of the next plane's taps made the kernel SLOWER by exactly its instruction
// cost, i.e. the kernel is bound by instruction issue. This generation halves the instruction count instead:
//
//   * a thread owns the voxel pair (z, z + 1) of one (x, y) row, so everything outside the gather -- the Laplacian of the
//     previous gradient (reference gradients.tpp:106-172), the data/Tikhonov combination (optimizer.tpp:191-200) and the
//     axis-0 filter chain (convolution.cpp:240-267) -- runs on packed f32x2 registers (FMUL2/FFMA2: IEEE round-to-
//     nearest per lane, no contraction, bit-identical to the scalar sequence) and every shared-memory operand is one
//     64-bit LDS;
//   * the Laplacian's border rules are applied per thread (z faces: one predicated instruction; y faces: a warp-uniform
//     branch; x faces: a block-uniform branch), no longer per block, so the tiles on the faces of the volume run the
//     packed path too;
//   * the filter taps and the packed constants live in the constant bank (kernel parameters), not in registers;
//   * slab decomposition (slab.py) is a template parameter of the same kernel: global plane numbers for the lookups and
//     the border rules, pack-region guard for the gather.
//
// Tiles, TMA ring and mbarrier protocol are those of kernels3d_tma.cuh (box = 64 z x TY rows of one plane).
#pragma once

#include "kernels3d_tma.cuh"

namespace lsf {

struct PairArgs {
	unsigned long long k2[7];  // flipped taps, each duplicated into both lanes: k2[q] multiplies in[i - R + q]
	unsigned long long one2;   // {1, 1}: opaque multiplier of the packed adds (see add2)
	unsigned long long neg2;   // {-1, -1}: a - b = fma(b, neg2, a)
	unsigned long long two2;   // {2, 2}
	unsigned long long strength2;
	int x_chunk;               // output planes per block along axis 0
	int x_lo, x_hi;            // allocation planes on which stage 1 may be evaluated: [x_lo, x_hi) (whole volume: 0, X)
	int x_fetch_hi;            // planes [.., x_fetch_hi) exist in the allocation (slab: x_hi + 1 unless x_hi is the volume's end)
};

#ifdef __CUDACC__

__device__ __forceinline__ f32x2 lds_f32x2(uint32_t addr) {
	f32x2 v;
	asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b, f32x2 neg) {
	f32x2 r;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(b), "l"(neg), "l"(a));
	return r;
}

template<bool TIKHONOV, int TY>
struct PairTile {
	static constexpr int TZ = 64;                     // voxels per block along z (32 lanes x 2)
	static constexpr int GZ = TZ + 8, GY = TY + 2;    // g_prev box: 1-voxel halo; along z the box starts 4 voxels early
	static constexpr int GZ0 = 4;                     // (a TMA box must start on a 16-byte boundary)
	static constexpr int GP_TX = TIKHONOV ? 3 * GY * GZ * 4 : 0;
	static constexpr int GP_BYTES = (GP_TX + 127) / 128 * 128;
	static constexpr int WP_BYTES = 3 * TY * TZ * 4;
	static constexpr int CN_BYTES = TY * TZ * 4;
	static constexpr int STAGE_BYTES = GP_BYTES + WP_BYTES + CN_BYTES;
	static constexpr int STAGE_TX = GP_TX + WP_BYTES + CN_BYTES;
};

// Stage 1 (gather + data term + Tikhonov term, reference optimizer.tpp:186-200) and the axis-0 filter pass
// (convolution.cpp:240-267) for the voxel pairs of a 64 x TY tile, marching along axis 0 over x_chunk output planes.
// Requires Z % 64 == 0 and Y % TY == 0 (the host falls back to k_hier_stage1_tma otherwise).
template<bool TIKHONOV, int R, int NS, int TY, bool SLAB>
static __global__ void __launch_bounds__(32 * TY, 512 / (32 * TY)) k_hier_stage1_pair(
		const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_w,
		const __grid_constant__ CUtensorMap map_c, const __grid_constant__ HierIterArgs a,
		const __grid_constant__ PairArgs t) {
	typedef PairTile<TIKHONOV, TY> T;
	constexpr int K = 2 * R + 1;
	if (a.check_convergence && level_converged(a.max_sq_bits, a.iteration, a.threshold)) return;
	extern __shared__ __align__(128) unsigned char stage_memory[];
	__shared__ uint64_t full[NS];
	__shared__ uint64_t empty[NS];  // one arrival per warp once it has read the slot

	const int Y = a.g.Y, Z = a.g.Z;
	const int tz = threadIdx.x, ty = threadIdx.y;
	const int z0 = blockIdx.x * T::TZ, y0 = blockIdx.y * TY;
	const int z = z0 + 2 * tz, y = y0 + ty;
	const int YZ = Y * Z;
	const int N = (int) a.g.N;
	const int xs = a.x_begin + blockIdx.z * t.x_chunk;
	const int xe = min(a.x_end, xs + t.x_chunk);
	const int x_first = max(xs - R, t.x_lo);   // planes outside the volume contribute zeros to the filter
	const int x_stop = xe + R;
	const int x_limit = min(x_stop, t.x_hi);   // stage 1 is evaluated on planes [x_first, x_limit)
	const int p_last = min(x_limit, t.x_fetch_hi - 1);  // last plane fetched (the Tikhonov term looks one plane ahead)
	const bool leader = tz == 0 && ty == 0;

	if (leader) {
#pragma unroll
		for (int s = 0; s < NS; s++) {
			mbar_init(&full[s], 1);
			mbar_init(&empty[s], TY);
		}
		mbar_fence_init();
	}
	__syncthreads();
	auto fetch = [&](int plane, int slot) {
		unsigned char* dst = stage_memory + slot * T::STAGE_BYTES;
		mbar_expect_tx(&full[slot], T::STAGE_TX);
		if (TIKHONOV) tma_load_4d(dst, &map_g, z0 - T::GZ0, y0 - 1, plane, 0, &full[slot]);
		tma_load_4d(dst + T::GP_BYTES, &map_w, z0, y0, plane, 0, &full[slot]);
		tma_load_4d(dst + T::GP_BYTES + T::WP_BYTES, &map_c, z0, y0, plane, 0, &full[slot]);
	};
	if (leader) {
		for (int s = 0; s < NS; s++)
			if (x_first + s <= p_last) fetch(x_first + s, s);
	}

	const uint32_t stage_base = smem_addr(stage_memory);
	const uint32_t off_g = ((ty + 1) * T::GZ + 2 * tz + T::GZ0) * 4;  // centre pair of component 0 in the halo'd tile
	const uint32_t off_w = T::GP_BYTES + (ty * T::TZ + 2 * tz) * 4;
	const uint32_t off_c = T::GP_BYTES + T::WP_BYTES + (ty * T::TZ + 2 * tz) * 4;
	constexpr uint32_t G_COMP = T::GY * T::GZ * 4, W_COMP = TY * T::TZ * 4, G_ROW = T::GZ * 4;

	const f32x2 one = t.one2, neg = t.neg2;
	f32x2 acc[3][K];
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int q = 0; q < K; q++) acc[c][q] = 0ull;

	// Laplacian border rules (reference gradients.tpp:28-35): z faces per thread, y faces per warp (a warp is one row)
	const bool z_first = z == 0, z_last = z + 2 == Z;
	const bool y_face = y == 0 || y == Y - 1;
	const int kind_y = border_kind(y, Y);
	const int kind_z_lo = border_kind(z, Z), kind_z_hi = border_kind(z + 1, Z);
	const int Xg = SLAB ? a.X_global : a.g.X;
	const int origin = SLAB ? a.x_origin : 0;
	const int sy = Z + 4, sx = (Y + 4) * (Z + 4);
	const float fy = (float) y, fz_lo = (float) z, fz_hi = (float) (z + 1);

	int slot = 0;
	uint32_t phase = 0;
	f32x2 prev[3] = { 0ull, 0ull, 0ull }, cur[3] = { 0ull, 0ull, 0ull };
	if (TIKHONOV) {
		mbar_wait(&full[0], 0);
#pragma unroll
		for (int c = 0; c < 3; c++) {
			cur[c] = lds_f32x2(stage_base + off_g + c * G_COMP);
			if (x_first + origin > 0) {
				const float2 v = __ldg(reinterpret_cast<const float2*>(a.g_prev + c * N + (x_first - 1) * YZ + y * Z + z));
				prev[c] = pack2(v.x, v.y);
			}
		}
	}
	int out = (x_first - R) * YZ + y * Z + z;  // index of the plane completed by the current step
#pragma unroll 1
	for (int x = x_first; x < x_stop; x++, out += YZ) {
		f32x2 g[3] = { 0ull, 0ull, 0ull };
		if (x < x_limit) {
			const uint32_t st = stage_base + slot * T::STAGE_BYTES;
			mbar_wait(&full[slot], phase);
			const f32x2 w0 = lds_f32x2(st + off_w), w1 = lds_f32x2(st + off_w + W_COMP), w2 = lds_f32x2(st + off_w + 2 * W_COMP);
			const f32x2 cn2 = lds_f32x2(st + off_c);
			f32x2 lap[3] = { 0ull, 0ull, 0ull };
			if (TIKHONOV) {
				const int xg = x + origin;
				const bool has_next = xg + 1 < Xg;
				const int next_slot = slot + 1 == NS ? 0 : slot + 1;
				if (has_next) mbar_wait(&full[next_slot], next_slot == 0 ? phase ^ 1u : phase);
				const uint32_t nst = stage_base + next_slot * T::STAGE_BYTES;
				if (xg == 0 || !has_next || y_face) {
					// faces of the volume along x or y: per-lane selects
					const int kind_x = border_kind(xg, Xg);
#pragma unroll
					for (int c = 0; c < 3; c++) {
						const uint32_t p = st + off_g + c * G_COMP;
						float n_lo = 0.0f, n_hi = 0.0f, p_lo, p_hi, c_lo, c_hi, ym_lo, ym_hi, yp_lo, yp_hi;
						if (has_next) unpack2(lds_f32x2(nst + off_g + c * G_COMP), n_lo, n_hi);
						unpack2(prev[c], p_lo, p_hi);
						unpack2(cur[c], c_lo, c_hi);
						unpack2(lds_f32x2(p - G_ROW), ym_lo, ym_hi);
						unpack2(lds_f32x2(p + G_ROW), yp_lo, yp_hi);
						const float zm = lds_f32(p - 4), zp = lds_f32(p + 8);
						float l_lo = laplace_select(p_lo, c_lo, n_lo, kind_x);
						l_lo += laplace_select(ym_lo, c_lo, yp_lo, kind_y);
						l_lo += laplace_select(zm, c_lo, c_hi, kind_z_lo);
						float l_hi = laplace_select(p_hi, c_hi, n_hi, kind_x);
						l_hi += laplace_select(ym_hi, c_hi, yp_hi, kind_y);
						l_hi += laplace_select(c_lo, c_hi, zp, kind_z_hi);
						lap[c] = pack2(l_lo, l_hi);
						prev[c] = cur[c];
						cur[c] = pack2(n_lo, n_hi);
					}
				} else {
#pragma unroll
					for (int c = 0; c < 3; c++) {
						const uint32_t p = st + off_g + c * G_COMP;
						const f32x2 next = lds_f32x2(nst + off_g + c * G_COMP);
						const f32x2 ym = lds_f32x2(p - G_ROW), yp = lds_f32x2(p + G_ROW);
						const float zm = lds_f32(p - 4), zp = lds_f32(p + 8);
						const f32x2 twice = mul2(cur[c], t.two2);
						f32x2 l = add2(sub2(next, twice, neg), prev[c], one);
						l = add2(l, add2(sub2(yp, twice, neg), ym, one), one);
						float c_lo, c_hi, t_lo, t_hi;
						unpack2(cur[c], c_lo, c_hi);
						unpack2(twice, t_lo, t_hi);
						float zt_lo = (c_hi - t_lo) + zm;
						float zt_hi = (zp - t_hi) + c_lo;
						if (z_first) zt_lo = c_hi - c_lo;
						if (z_last) zt_hi = c_lo - c_hi;
						lap[c] = add2(l, pack2(zt_lo, zt_hi), one);
						prev[c] = cur[c];
						cur[c] = next;
					}
				}
			}
			// this warp (one row of the tile) has read everything it needs from `slot`; the leader refills the slot of the
			// previous plane as soon as all warps have released it -- no block-wide barrier, the warps of a block drift up
			// to NS - 1 planes apart and their gather latencies overlap
			__syncwarp();
			if (tz == 0) mbar_arrive(&empty[slot]);
			if (leader && x > x_first) {
				const int refill = slot == 0 ? NS - 1 : slot - 1;
				mbar_wait(&empty[refill], slot == 0 ? phase ^ 1u : phase);
				if (x - 1 + NS <= p_last) fetch(x - 1 + NS, refill);
			}
			float wx_lo, wx_hi, wy_lo, wy_hi, wz_lo, wz_hi, cn_lo, cn_hi;
			unpack2(w0, wx_lo, wx_hi);
			unpack2(w1, wy_lo, wy_hi);
			unpack2(w2, wz_lo, wz_hi);
			unpack2(cn2, cn_lo, cn_hi);
			const float fx = (float) (x + origin);
			const float4 s_lo = gather4q<SLAB>(a, sx, sy, fx, fy, fz_lo, wx_lo, wy_lo, wz_lo, one);
			const float4 s_hi = gather4q<SLAB>(a, sx, sy, fx, fy, fz_hi, wx_hi, wy_hi, wz_hi, one);
			const float d_lo = s_lo.x - cn_lo, d_hi = s_hi.x - cn_hi;
			g[0] = pack2((s_lo.y * d_lo) * a.amplifier, (s_hi.y * d_hi) * a.amplifier);
			g[1] = pack2((s_lo.z * d_lo) * a.amplifier, (s_hi.z * d_hi) * a.amplifier);
			g[2] = pack2((s_lo.w * d_lo) * a.amplifier, (s_hi.w * d_hi) * a.amplifier);
			if (TIKHONOV) {
#pragma unroll
				for (int c = 0; c < 3; c++) g[c] = sub2(g[c], mul2(lap[c], t.strength2), neg);
			}
		}
		// axis-0 filter pass: plane x is tap q of output plane x + R - q; acc[c][q] holds the partial sum (taps 0..q) of
		// output x + R - q, so adding in place from the oldest output down reproduces sum_{q ascending} in[.]*k[q]
#pragma unroll
		for (int c = 0; c < 3; c++) {
#pragma unroll
			for (int q = K - 1; q >= 1; q--) acc[c][q] = add2(acc[c][q - 1], mul2(g[c], t.k2[q]), one);
			acc[c][0] = mul2(g[c], t.k2[0]);
		}
		if (x - R >= xs) {
#pragma unroll
			for (int c = 0; c < 3; c++) {
				float2 v;
				unpack2(acc[c][K - 1], v.x, v.y);
				*reinterpret_cast<float2*>(a.g_out + c * N + out) = v;
			}
		}
		slot++;
		if (slot == NS) {
			slot = 0;
			phase ^= 1u;
		}
	}
}

inline bool pair_supported(const Grid3& g, int ty) {
	return g.Z % 64 == 0 && g.Y % ty == 0;
}

inline unsigned long long dup2(float v) {
	unsigned bits;
	memcpy(&bits, &v, 4);
	return ((unsigned long long) bits << 32) | bits;
}

template<bool TIKHONOV, int TY>
int ensure_pair_maps(TmaMaps& maps, const Grid3& g, const float* warp, const float* canonical, const float* g_prev) {
	typedef PairTile<TIKHONOV, TY> T;
	if (maps.key[0] == warp && maps.key[1] == canonical && maps.key[2] == g_prev && maps.tile_y == TY + 1000) return LSF_OK;
	LSF_TRY(make_planes_map(&maps.warp, warp, 3, g.N, g, T::TZ, TY));
	LSF_TRY(make_planes_map(&maps.canonical, canonical, 1, g.N, g, T::TZ, TY));
	LSF_TRY(make_planes_map(&maps.g_prev, g_prev, 3, g.N, g, T::GZ, T::GY));
	maps.key[0] = warp;
	maps.key[1] = canonical;
	maps.key[2] = g_prev;
	maps.tile_y = TY + 1000;
	return LSF_OK;
}

// x_lo / x_hi: allocation planes on which stage 1 may be evaluated (whole volume: 0 and X)
template<bool TIKHONOV, int R, int TY, bool SLAB>
int launch_stage1_pair(TmaMaps& maps, HierIterArgs a, const Taps& taps, float* h, int x_chunk, int x_lo, int x_hi,
		int x_fetch_hi, cudaStream_t stream) {
	typedef PairTile<TIKHONOV, TY> T;
	constexpr int NS = 4;
	LSF_TRY((ensure_pair_maps<TIKHONOV, TY>(maps, a.g, a.warp, a.canonical, a.g_prev)));
	PairArgs t;
	for (int q = 0; q < 7; q++) t.k2[q] = dup2(q < 2 * R + 1 ? taps.k[q] : 0.0f);
	t.one2 = dup2(1.0f);
	t.neg2 = dup2(-1.0f);
	t.two2 = dup2(2.0f);
	t.strength2 = dup2(a.strength);
	t.x_chunk = x_chunk;
	t.x_lo = x_lo;
	t.x_hi = x_hi;
	t.x_fetch_hi = x_fetch_hi;
	a.g_out = h;
	const dim3 block(32, TY, 1), grid(a.g.Z / T::TZ, a.g.Y / TY, div_up(a.x_end - a.x_begin, x_chunk));
	const size_t shared = (size_t) NS * T::STAGE_BYTES;
	static bool configured = false;
	if (!configured) {
		LSF_CUDA(cudaFuncSetAttribute(k_hier_stage1_pair<TIKHONOV, R, NS, TY, SLAB>,
				cudaFuncAttributeMaxDynamicSharedMemorySize, (int) shared));
		configured = true;
	}
	k_hier_stage1_pair<TIKHONOV, R, NS, TY, SLAB> <<<counted(grid), block, shared, stream>>>(maps.g_prev, maps.warp,
			maps.canonical, a, t);
	return LSF_OK;
}

// ---------------------------------------------------------------------------------------------- axis-1 / axis-2 passes, paired
// k_sobolev_ymarch with two z voxels per thread (same passes, same order of operations: convolution.cpp:268-331,
// optimizer.tpp:207-211). The axis-1 chain, the warp update and the squared norm run on packed f32x2 registers. For
// the axis-2 pass the row is kept twice in shared memory, once as is (A) and once shifted by one column (B[i] =
// A[i + 1]): every tap's operand pair (v[z - R + q], v[z + 1 - R + q]) is then one aligned 64-bit LDS from A or B.
struct YMarch2Args {
	const float* in;   // planes after the axis-0 pass
	float* out;        // planes: filtered gradient (nullptr when nothing reads it)
	float* warp;       // planes, updated in place: warp -= out * rate (nullptr: deferred to the next stage 1)
	Grid3 g;
	unsigned long long k2[7];  // flipped taps duplicated into both lanes
	unsigned long long one2, neg2, rate2;
	float threshold;
	unsigned* max_sq_bits;
	int iteration;
	int check_convergence;
	int y_chunk;       // output rows per block
	int x_begin;       // first plane (blockIdx.y counts from here)
	int tile_z;        // output columns per block (even); blockDim.x = tile_z / 2 (+ 32 halo threads when gridDim.x > 1)
	int batch_X;       // batch of pairs (HierIterArgs::batch_X): planes per pair, 0 = one volume
	int batch_slot_stride;  // convergence slots per pair
};

#ifdef __CUDACC__
__device__ __forceinline__ void sts_f32x2(uint32_t addr, f32x2 v) {
	asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
	asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

template<int R>
static __global__ void __launch_bounds__(288) k_sobolev_ymarch2(const __grid_constant__ YMarch2Args a) {
	constexpr int K = 2 * R + 1;
	constexpr int H = 4;  // halo columns kept either side of the tile (>= R, even)
	unsigned* slots = a.max_sq_bits;
	if (a.batch_X > 0 && slots != nullptr) slots += ((a.x_begin + blockIdx.y) / a.batch_X) * a.batch_slot_stride;
	if (a.check_convergence && level_converged(slots, a.iteration, a.threshold)) return;
	extern __shared__ __align__(16) float row_memory2[];  // [2 buffers][3 components][A: W | B: W]
	const int Y = a.g.Y, Z = a.g.Z;
	const int NT = a.tile_z / 2;   // owner threads
	const int W = a.tile_z + 2 * H;
	const int tid = threadIdx.x;
	const int z0 = blockIdx.x * a.tile_z;
	const int x = a.x_begin + blockIdx.y;
	const int ys = blockIdx.z * a.y_chunk;
	const int ye = min(Y, ys + a.y_chunk);
	// column pair of this thread in the row buffer: owners hold H .. H + tile_z - 1, the halo threads H columns either side
	const bool owner = tid < NT;
	const int j = tid - NT;
	const int il = owner ? H + 2 * tid : (j < H / 2 ? 2 * j : a.tile_z + H + 2 * (j - H / 2));
	const int z = z0 - H + il;
	const bool active = (owner || j < H) && z >= 0 && z < Z;  // Z is even: a pair is inside or outside as a whole
	const bool writes = owner && active;
	for (int i = tid; i < 12 * W; i += blockDim.x) row_memory2[i] = 0.0f;  // columns outside the volume stay zero
	__syncthreads();
	const uint32_t rows = smem_addr(row_memory2);
	const f32x2 one = a.one2, neg = a.neg2;

	f32x2 acc[3][K];
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int q = 0; q < K; q++) acc[c][q] = 0ull;
	const int r_first = max(ys - R, 0);
	const int r_stop = ye + R;
	const int r_load_end = min(r_stop, Y);
	const int N = (int) a.g.N;
	int at = (x * Y + r_first) * Z + z;  // row being consumed
	f32x2 next[3] = { 0ull, 0ull, 0ull };
	if (active) {
#pragma unroll
		for (int c = 0; c < 3; c++) {
			const float2 v = __ldg(reinterpret_cast<const float2*>(a.in + c * N + at));
			next[c] = pack2(v.x, v.y);
		}
	}
	float best = 0.0f;
	int buffer = 0;
#pragma unroll 1
	for (int r = r_first; r < r_stop; r++, at += Z) {
		const f32x2 v0 = next[0], v1 = next[1], v2 = next[2];
		if (active && r + 1 < r_load_end) {
#pragma unroll
			for (int c = 0; c < 3; c++) {
				const float2 v = __ldg(reinterpret_cast<const float2*>(a.in + c * N + at + Z));
				next[c] = pack2(v.x, v.y);
			}
		} else {
			next[0] = next[1] = next[2] = 0ull;
		}
		// axis-1 pass: row r is tap q of output row r + R - q (same chain as the axis-0 pass of stage 1)
#pragma unroll
		for (int q = K - 1; q >= 1; q--) {
			acc[0][q] = add2(acc[0][q - 1], mul2(v0, a.k2[q]), one);
			acc[1][q] = add2(acc[1][q - 1], mul2(v1, a.k2[q]), one);
			acc[2][q] = add2(acc[2][q - 1], mul2(v2, a.k2[q]), one);
		}
		acc[0][0] = mul2(v0, a.k2[0]);
		acc[1][0] = mul2(v1, a.k2[0]);
		acc[2][0] = mul2(v2, a.k2[0]);
		if (r - R < ys) continue;  // block-uniform: still priming
		const uint32_t row = rows + buffer * (6 * W * 4);
		f32x2 w[3] = { 0ull, 0ull, 0ull };
		const int o = at - R * Z;  // voxel (x, r - R, z)
		if (active) {
#pragma unroll
			for (int c = 0; c < 3; c++) {
				const uint32_t pa = row + (c * 2 * W + il) * 4;  // A[il], A[il + 1]
				float lo, hi;
				unpack2(acc[c][K - 1], lo, hi);
				sts_f32x2(pa, acc[c][K - 1]);
				if (il > 0) sts_f32(pa + (W - 1) * 4, lo);  // B[il - 1] = A[il]
				sts_f32(pa + W * 4, hi);        // B[il] = A[il + 1]
			}
		}
		if (writes && a.warp != nullptr) {
#pragma unroll
			for (int c = 0; c < 3; c++) {
				const float2 v = *reinterpret_cast<const float2*>(a.warp + c * N + o);
				w[c] = pack2(v.x, v.y);
			}
		}
		__syncthreads();
		if (writes) {
			// axis-2 pass: tap q multiplies the pair (A[il - R + q], A[il - R + q + 1])
			f32x2 gq[3];
#pragma unroll
			for (int c = 0; c < 3; c++) {
				const uint32_t pa = row + (c * 2 * W + il - R) * 4;
				f32x2 sum = 0ull;
#pragma unroll
				for (int q = 0; q < K; q++) {
					// il is even: the pair starts on an even column when (q - R) is even, else take it from the shifted copy
					const f32x2 v = ((q - R) % 2 == 0) ? lds_f32x2(pa + q * 4) : lds_f32x2(pa + (W + q - 1) * 4);
					sum = q == 0 ? mul2(v, a.k2[0]) : add2(sum, mul2(v, a.k2[q]), one);
				}
				gq[c] = sum;
			}
#pragma unroll
			for (int c = 0; c < 3; c++) {
				float2 v;
				if (a.out != nullptr) {
					unpack2(gq[c], v.x, v.y);
					*reinterpret_cast<float2*>(a.out + c * N + o) = v;
				}
				if (a.warp != nullptr) {
					unpack2(sub2(w[c], mul2(gq[c], a.rate2), neg), v.x, v.y);
					*reinterpret_cast<float2*>(a.warp + c * N + o) = v;
				}
			}
			f32x2 sq = mul2(gq[0], gq[0]);
			sq = add2(sq, mul2(gq[1], gq[1]), one);
			sq = add2(sq, mul2(gq[2], gq[2]), one);
			float sq_lo, sq_hi;
			unpack2(sq, sq_lo, sq_hi);
			if (sq_lo > best) best = sq_lo;
			if (sq_hi > best) best = sq_hi;
		}
		buffer ^= 1;  // the row written two steps from now is read by nobody after the next barrier
	}
	if (slots != nullptr) block_atomic_max(best, slots + a.iteration);
}

inline bool ymarch2_supported(const Grid3& g, const float* h, const float* filtered, const float* warp) {
	auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; };
	return g.Z % 2 == 0 && g.N % 2 == 0 && aligned(h) && aligned(filtered) && aligned(warp);
}

template<int R>
void launch_ymarch2(const Taps& taps, const HierIterArgs& a, const float* h, float* filtered, float* warp, int y_chunk,
		cudaStream_t stream, int x_begin = 0, int planes = -1) {
	const Grid3& g = a.g;
	YMarch2Args f;
	f.in = h;
	f.out = filtered;
	f.warp = warp;
	f.g = g;
	for (int q = 0; q < 7; q++) f.k2[q] = dup2(q < 2 * R + 1 ? taps.k[q] : 0.0f);
	f.one2 = dup2(1.0f);
	f.neg2 = dup2(-1.0f);
	f.rate2 = dup2(a.rate);
	f.threshold = a.threshold;
	f.max_sq_bits = a.max_sq_bits;
	f.iteration = a.iteration;
	f.check_convergence = a.check_convergence;
	f.y_chunk = y_chunk;
	f.x_begin = x_begin;
	f.batch_X = a.batch_X;
	f.batch_slot_stride = a.batch_slot_stride;
	f.tile_z = std::min(512, (int) div_up(g.Z, 64) * 64);
	const int tiles = div_up(g.Z, f.tile_z);
	const dim3 grid(tiles, planes < 0 ? g.X : planes, div_up(g.Y, y_chunk));
	const int threads = f.tile_z / 2 + (tiles > 1 ? 32 : 0);
	const size_t shared = (size_t) 12 * (f.tile_z + 8) * sizeof(float);
	k_sobolev_ymarch2<R> <<<counted(grid), threads, shared, stream>>>(f);
}
#endif  // __CUDACC__

// best available kernel for the axis-1 / axis-2 passes: k_sobolev_ymarch3 (kernels3d_ymarch3.cuh) where it applies, else
// k_sobolev_ymarch2 (defined in kernels3d_ymarch3.cuh, which includes this header)
#ifdef __CUDACC__
template<int R>
void launch_ymarch_auto(const Taps& taps, const HierIterArgs& a, const float* h, float* filtered, float* warp, int y_chunk,
		cudaStream_t stream, int x_begin = 0, int planes = -1);
#endif

// ---------------------------------------------------------------------------------------------- axis-0 pass alone
// Slab mode (slab.py) exchanges the halo planes of the unfiltered gradient between stage 1 and the filter, so the
// axis-0 pass cannot ride on stage 1 there: this kernel does it on its own (reference convolution.cpp:240-267), a thread
// marching along axis 0 with one z pair. Planes outside the allocation count as zeros (the allocation's halo planes
// beyond the volume are zero as well).
struct XMarchArgs {
	const float* in;
	float* out;
	int X, Y, Z;           // allocation
	int x_begin, x_end;    // output planes
	unsigned long long k2[7];
	unsigned long long one2;
	unsigned* max_sq_bits;
	float threshold;
	int iteration, check_convergence;
	int chunk;
};

#ifdef __CUDACC__
template<int R>
static __global__ void __launch_bounds__(256) k_hier_xmarch(const __grid_constant__ XMarchArgs a) {
	constexpr int K = 2 * R + 1;
	if (a.check_convergence && level_converged(a.max_sq_bits, a.iteration, a.threshold)) return;
	const int pair = blockIdx.x * blockDim.x + threadIdx.x;
	if (pair >= a.Y * a.Z / 2) return;
	const int YZ = a.Y * a.Z;
	const int N = a.X * YZ;
	const int xs = a.x_begin + blockIdx.y * a.chunk;
	const int xe = min(a.x_end, xs + a.chunk);
	const int x_first = max(xs - R, 0), x_stop = xe + R;
	const f32x2 one = a.one2;
	f32x2 acc[3][K];
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int q = 0; q < K; q++) acc[c][q] = 0ull;
	int at = x_first * YZ + 2 * pair;
	f32x2 next[3];
#pragma unroll
	for (int c = 0; c < 3; c++) {
		const float2 v = __ldg(reinterpret_cast<const float2*>(a.in + c * N + at));
		next[c] = pack2(v.x, v.y);
	}
#pragma unroll 1
	for (int x = x_first; x < x_stop; x++, at += YZ) {
		const f32x2 v0 = next[0], v1 = next[1], v2 = next[2];
		if (x + 1 < a.X && x + 1 < x_stop) {
#pragma unroll
			for (int c = 0; c < 3; c++) {
				const float2 v = __ldg(reinterpret_cast<const float2*>(a.in + c * N + at + YZ));
				next[c] = pack2(v.x, v.y);
			}
		} else {
			next[0] = next[1] = next[2] = 0ull;
		}
#pragma unroll
		for (int q = K - 1; q >= 1; q--) {
			acc[0][q] = add2(acc[0][q - 1], mul2(v0, a.k2[q]), one);
			acc[1][q] = add2(acc[1][q - 1], mul2(v1, a.k2[q]), one);
			acc[2][q] = add2(acc[2][q - 1], mul2(v2, a.k2[q]), one);
		}
		acc[0][0] = mul2(v0, a.k2[0]);
		acc[1][0] = mul2(v1, a.k2[0]);
		acc[2][0] = mul2(v2, a.k2[0]);
		if (x - R >= xs) {
			const int o = at - R * YZ;
#pragma unroll
			for (int c = 0; c < 3; c++) {
				float2 v;
				unpack2(acc[c][K - 1], v.x, v.y);
				*reinterpret_cast<float2*>(a.out + c * N + o) = v;
			}
		}
	}
}

// Filter phase of a slab iteration: axis-0 pass of planes [x_begin, x_end) into `h`, then the paired y-marching kernel
// (axes 1 and 2, warp update, max-norm).
template<int R>
void launch_slab_filter(const Taps& taps, const HierIterArgs& a, const float* in, float* h, float* filtered, float* warp,
		cudaStream_t stream) {
	XMarchArgs f;
	f.in = in;
	f.out = h;
	f.X = a.g.X;
	f.Y = a.g.Y;
	f.Z = a.g.Z;
	f.x_begin = a.x_begin;
	f.x_end = a.x_end;
	for (int q = 0; q < 7; q++) f.k2[q] = dup2(q < 2 * R + 1 ? taps.k[q] : 0.0f);
	f.one2 = dup2(1.0f);
	f.max_sq_bits = a.max_sq_bits;
	f.threshold = a.threshold;
	f.iteration = a.iteration;
	f.check_convergence = a.check_convergence;
	const int planes = a.x_end - a.x_begin;
	const int plane_blocks = (int) div_up((long long) f.Y * f.Z / 2, 256);
	f.chunk = marching_chunk(planes, plane_blocks, 2 * R, 3);
	k_hier_xmarch<R> <<<counted(dim3(plane_blocks, (unsigned) div_up(planes, f.chunk))), 256, 0, stream>>>(f);
	const int tiles = (int) div_up(a.g.Z, 512);
	const int y_chunk = marching_chunk(a.g.Y, tiles * planes, 2 * R, 6);
	launch_ymarch_auto<R>(taps, a, h, filtered, warp, y_chunk, stream, a.x_begin, planes);
}
#endif  // __CUDACC__

// One whole-volume iteration: paired stage 1 (tile_y = 4 or 8) when the level's shape allows it, else the third
// generation; then the y-marching filter. Returns a negative status on failure.
// Pending update of a deferred iteration (see k_hier_stage1_tma APPLY): warp_out = warp - g * rate
static __global__ void k_apply_update3d(const float* __restrict__ warp, const float* __restrict__ g, float* __restrict__ warp_out,
		float rate, long long count) {
	const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count) warp_out[i] = warp[i] - g[i] * rate;
}

inline bool deferred_update_supported(const Grid3& g, const float* h, const float* filtered) {
	return ymarch2_supported(g, h, filtered, filtered);
}

// One iteration with the warp update deferred into the next stage 1: a.warp (read, before the previous update),
// a.warp_out (written, after it) and a.g_prev = filtered (previous gradient in, this iteration's gradient out).
template<int R>
int launch_iteration_deferred(TmaMaps& maps, HierIterArgs a, const Taps& taps, float* h, float* filtered, int x_chunk,
		int y_chunk, cudaStream_t stream, cudaEvent_t* events) {
	LSF_TRY((launch_stage1_tma<true, R, true, true>(maps, a, taps, h, x_chunk, stream)));
	if (events) cudaEventRecord(events[1], stream);
	launch_ymarch_auto<R>(taps, a, h, filtered, nullptr, y_chunk, stream);
	if (events) cudaEventRecord(events[2], stream);
	return LSF_OK;
}

template<int R>
int launch_iteration_v4(bool tikhonov, TmaMaps& maps, HierIterArgs a, const Taps& taps, float* h, float* filtered,
		float* warp, int x_chunk, int y_chunk, int tile_y, cudaStream_t stream, cudaEvent_t* events) {
	const int X = a.g.X;
	if ((tile_y != 4 && tile_y != 8) || !pair_supported(a.g, tile_y)) {
		const bool coupled = getenv("LSF_DECOUPLE") && getenv("LSF_DECOUPLE")[0] == '0';  // A/B: block barrier per plane
		if (tikhonov && coupled) LSF_TRY((launch_stage1_tma<true, R, false>(maps, a, taps, h, x_chunk, stream)));
		else if (tikhonov) LSF_TRY((launch_stage1_tma<true, R>(maps, a, taps, h, x_chunk, stream)));
		else LSF_TRY((launch_stage1_tma<false, R>(maps, a, taps, h, x_chunk, stream)));
	} else if (tikhonov && tile_y == 8) LSF_TRY((launch_stage1_pair<true, R, 8, false>(maps, a, taps, h, x_chunk, 0, X, X, stream)));
	else if (tikhonov) LSF_TRY((launch_stage1_pair<true, R, 4, false>(maps, a, taps, h, x_chunk, 0, X, X, stream)));
	else if (tile_y == 8) LSF_TRY((launch_stage1_pair<false, R, 8, false>(maps, a, taps, h, x_chunk, 0, X, X, stream)));
	else LSF_TRY((launch_stage1_pair<false, R, 4, false>(maps, a, taps, h, x_chunk, 0, X, X, stream)));
	if (events) cudaEventRecord(events[1], stream);
	const bool scalar_filter = getenv("LSF_YMARCH2") && getenv("LSF_YMARCH2")[0] == '0';  // A/B: one voxel per thread
	if (!scalar_filter && ymarch2_supported(a.g, h, filtered, warp)) launch_ymarch_auto<R>(taps, a, h, filtered, warp, y_chunk, stream);
	else launch_ymarch<R>(taps, a, h, filtered, warp, y_chunk, stream);
	if (events) cudaEventRecord(events[2], stream);
	return LSF_OK;
}

#endif  // __CUDACC__

}  // namespace lsf
