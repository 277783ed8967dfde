// kernels3d_ymarch3.cuh -- fifth generation of the axis-1 / axis-2 filter passes of the 3D hierarchical iteration.
//
// SASS audit of k_sobolev_ymarch2<3> (profiles/r2_sass_audit.md): 240 issued instructions per thread and row for 123 that
// do the work -- 64-bit address arithmetic for every global access (LEA / LEA.HI.X / IADD3.X chains per component),
// register rotation MOVs of the prefetched row, shared-memory addresses re-derived from the toggling buffer index,
// CS2R zeroing, a three-way run-time branch over the store variants. The issue slots of that kernel are 54 % busy
// with 6 warps per scheduler, i.e. the instruction stream is what a row costs. This generation keeps the algorithm
// (thread = z pair of one x plane marching along y; axis-1 pass as a register chain of partial sums, reference
// convolution.cpp:268-299; axis-2 pass from a double-buffered shared row kept as is and shifted by one column,
// convolution.cpp:300-331; warp update and max-norm, optimizer.tpp:207-211) and removes the overhead:
//
//   * component planes arrive as three pointers in the kernel parameters; a thread addresses them with one 32-bit
//     float2 index (one IMAD.WIDE per access);
//   * the row loop is unrolled by two: the two shared-memory buffers and the two prefetch register sets are named at
//     compile time (no rotation, every LDS / STS is one base register + immediate), and every row is requested two rows
//     ahead of its use;
//   * the store variants (gradient only | gradient + warp update) are template parameters;
//   * the axis-2 pass reads the row as aligned pairs and builds the odd-aligned operand pairs in registers instead of
//     keeping a shifted copy of the row: 30 instead of 54 shared-memory wavefronts per warp and row -- ncu showed the
//     shared-memory pipe 70 % busy and the kernel bound by it (0.099 -> 0.083 ms per 256^3 iteration);
//   * symmetric kernels (every Sobolev kernel is: k[q] == k[K-1-q] bit for bit) multiply each input once per distinct
//     tap in the axis-1 chain: v * k[q] and v * k[K-1-q] are the same rounded product.
//
// Arithmetic per output is unchanged (float32, reference order, separate multiply and add), so results stay bit-identical.
#pragma once

#include "kernels3d_pair.cuh"

namespace lsf {

struct YMarch3Args {
	const float* in[3];   // component planes after the axis-0 pass
	float* out[3];        // component planes of the filtered gradient (HAS_OUT)
	float* warp[3];       // component planes of the warp, updated in place: warp -= out * rate (HAS_WARP)
	int X, Y, Z;
	unsigned long long k2[7];  // flipped taps duplicated into both lanes
	unsigned long long one2, neg2, rate2;
	float threshold;
	unsigned* max_sq_bits;
	int iteration;
	int check_convergence;
	int y_chunk;          // output rows per block
	int x_begin;          // first plane (blockIdx.y counts from here)
	int planes;           // balanced mode (y_chunk == 0): the planes x_begin .. x_begin + planes - 1 form one list of
	                      // planes * Y rows that the gridDim.y blocks split evenly (a block's range may span two planes)
	int batch_X;          // batch of pairs (HierIterArgs::batch_X): planes per pair, 0 = one volume
	int batch_slot_stride;  // convergence slots per pair
};

#ifdef __CUDACC__

template<int R, bool SYM, bool HAS_OUT, bool HAS_WARP, int TZ = 256>
static __global__ void __maxnreg__(HAS_WARP ? 96 : 80) k_sobolev_ymarch3(const __grid_constant__ YMarch3Args a) {
	constexpr int K = 2 * R + 1;
	constexpr int H = 4;             // halo columns kept either side of the tile (>= R, even)
	// TZ: output columns per block (256; 128 for rows of 128 .. 255 voxels: the 128^3 pyramid level)
	constexpr int W = TZ + 2 * H;    // columns of a shared row
	constexpr int NT = TZ / 2;       // owner threads
	constexpr uint32_t ROW_BYTES = W * 4;              // one component of one row
	constexpr uint32_t BUFFER_BYTES = 3 * ROW_BYTES;
	pdl_launch_dependents();
	__shared__ __align__(16) float row_memory3[2 * 3 * W];  // [2 buffers][3 components][W]
	const int Y = a.Y, Z = a.Z;
	const int tid = threadIdx.x;
	const int z0 = blockIdx.x * TZ;
	// column pair of this thread in the row buffer: owners hold H .. H + TZ - 1, the halo threads H columns either side
	const bool owner = tid < NT;
	const int j = tid - NT;
	const int il = owner ? H + 2 * tid : (j < H / 2 ? 2 * j : TZ + H + 2 * (j - H / 2));
	const int z = z0 - H + il;
	const bool active = (owner || j < H) && z >= 0 && z < Z;  // Z is even: a pair is inside or outside as a whole
	const bool writes = owner && active;
	for (int i = tid; i < 2 * 3 * W; i += blockDim.x) row_memory3[i] = 0.0f;  // columns outside the volume stay zero
	__syncthreads();
	pdl_wait();  // everything above is independent of the previous kernel's output
	const uint32_t mine = smem_addr(row_memory3) + il * 4;  // A[il] of component 0, buffer 0
	const f32x2 one = a.one2, neg = a.neg2;
	const float2* __restrict__ in0 = reinterpret_cast<const float2*>(a.in[0]);
	const float2* __restrict__ in1 = reinterpret_cast<const float2*>(a.in[1]);
	const float2* __restrict__ in2 = reinterpret_cast<const float2*>(a.in[2]);
	const int ZP = Z >> 1;  // float2 elements per row
	// this block's rows: one y-chunk of one plane (blockIdx.y = plane, blockIdx.z = chunk), or, in balanced mode, an
	// even share of the planes' rows taken as one list -- up to two segments (the tail of one plane, the head of the next)
	long long flat_lo = 0, flat_hi = 0;
	if (a.y_chunk == 0) {
		const long long total = (long long) a.planes * Y;
		flat_lo = total * blockIdx.y / gridDim.y;
		flat_hi = total * (blockIdx.y + 1) / gridDim.y;
	}
#pragma unroll 1
	for (int segment = 0; segment < 2; segment++) {
	int x, ys, ye;
	if (a.y_chunk > 0) {
		if (segment == 1) break;
		x = a.x_begin + blockIdx.y;
		ys = blockIdx.z * a.y_chunk;
		ye = min(Y, ys + a.y_chunk);
	} else {
		const int plane = (int) (flat_lo / Y) + segment;
		const long long plane_lo = (long long) plane * Y;
		if (flat_hi <= plane_lo || flat_lo >= flat_hi) break;
		x = a.x_begin + plane;
		ys = segment == 0 ? (int) (flat_lo - plane_lo) : 0;
		ye = (int) min((long long) Y, flat_hi - plane_lo);
	}
	unsigned* slots = a.max_sq_bits;
	if (a.batch_X > 0 && slots != nullptr) slots += (x / a.batch_X) * a.batch_slot_stride;
	if (a.check_convergence && level_converged(slots, a.iteration, a.threshold)) continue;

	f32x2 acc[3][K];
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int q = 0; q < K; q++) acc[c][q] = 0ull;
	const int r_first = max(ys - R, 0);
	const int r_stop = ye + R;
	const int r_load_end = min(r_stop, Y);
	// float2 index of the next row to request; threads without a column pair of their own keep re-reading element 0
	// (their values are never stored), so that the steady-state loads need no predicate
	int ld = active ? ((x * Y + r_first) * Z + z) >> 1 : 0;
	const int ld_step = active ? ZP : 0;
	int ld_row = r_first;
	int st = ((x * Y + ys) * Z + z) >> 1;                    // float2 index of the next output row

	auto request = [&](f32x2 (&v)[3]) {
		if (active && ld_row < r_load_end) {
			const float2 v0 = __ldg(in0 + ld), v1 = __ldg(in1 + ld), v2 = __ldg(in2 + ld);
			v[0] = pack2(v0.x, v0.y);
			v[1] = pack2(v1.x, v1.y);
			v[2] = pack2(v2.x, v2.y);
		} else {
			v[0] = v[1] = v[2] = 0ull;
		}
		ld += ld_step;
		ld_row++;
	};
	// the same for rows known to lie inside the volume
	auto request_inside = [&](f32x2 (&v)[3]) {
		const float2 v0 = __ldg(in0 + ld), v1 = __ldg(in1 + ld), v2 = __ldg(in2 + ld);
		v[0] = pack2(v0.x, v0.y);
		v[1] = pack2(v1.x, v1.y);
		v[2] = pack2(v2.x, v2.y);
		ld += ld_step;
		ld_row++;
	};
	// axis-1 pass: the row is tap q of output row (row + R - q); acc[c][q] holds the partial sum (taps 0..q) of that
	// output, so adding in place from the oldest output down reproduces sum_{q ascending} in[.] * k[q]
	auto chain = [&](const f32x2 (&v)[3]) {
#pragma unroll
		for (int c = 0; c < 3; c++) {
			f32x2 p[K];
#pragma unroll
			for (int q = 0; q < K; q++) p[q] = (SYM && q > R) ? p[K - 1 - q] : mul2(v[c], a.k2[q]);
#pragma unroll
			for (int q = K - 1; q >= 1; q--) acc[c][q] = add2(acc[c][q - 1], p[q], one);
			acc[c][0] = p[0];
		}
	};
	float best = 0.0f;
	// finishes the output row `st` from the chain's oldest partial sums through shared buffer BUF
	auto emit = [&](auto buffer_tag) {
		constexpr uint32_t BUF = decltype(buffer_tag)::value * BUFFER_BYTES;
		if (active) {
#pragma unroll
			for (int c = 0; c < 3; c++) {
				const uint32_t pa = mine + BUF + c * ROW_BYTES;  // A[il], A[il + 1]
				sts_f32x2(pa, acc[c][K - 1]);
			}
		}
		f32x2 w[3] = { 0ull, 0ull, 0ull };
		if (HAS_WARP && writes) {
#pragma unroll
			for (int c = 0; c < 3; c++) {
				const float2 v = reinterpret_cast<const float2*>(a.warp[c])[st];
				w[c] = pack2(v.x, v.y);
			}
		}
		__syncthreads();
		if (writes) {
			// axis-2 pass: tap q multiplies the pair (A[il - R + q], A[il - R + q + 1])
			f32x2 gq[3];
#pragma unroll
			for (int c = 0; c < 3; c++) {
				const uint32_t pa = mine + BUF + c * ROW_BYTES - R * 4;
				f32x2 sum = 0ull;
				// the K + 1 values A[il - R .. il + R + 1] arrive as aligned pairs (il is even; for an odd radius the first
				// pair starts one column early); operand pairs that start on an odd column are put together in registers.
				// A shifted copy of the row in shared memory (k_sobolev_ymarch2) saves those moves but costs 6 stores and
				// 12 more loads per row: the shared-memory pipe, not the ALU, bounds this kernel (profiles/r2_ncu_headline.md)
				constexpr int OFF = R % 2;
				constexpr int PAIRS = (K + 1 + 2 * OFF) / 2;
				float window[2 * PAIRS];
#pragma unroll
				for (int j = 0; j < PAIRS; j++) unpack2(lds_f32x2(pa + (2 * j - OFF) * 4), window[2 * j], window[2 * j + 1]);
#pragma unroll
				for (int q = 0; q < K; q++) {
					const f32x2 v = pack2(window[q + OFF], window[q + OFF + 1]);  // (A[il - R + q], A[il - R + q + 1])
					sum = q == 0 ? mul2(v, a.k2[0]) : add2(sum, mul2(v, a.k2[q]), one);
				}
				gq[c] = sum;
			}
#pragma unroll
			for (int c = 0; c < 3; c++) {
				float2 v;
				if (HAS_OUT) {
					unpack2(gq[c], v.x, v.y);
					reinterpret_cast<float2*>(a.out[c])[st] = v;
				}
				if (HAS_WARP) {
					unpack2(sub2(w[c], mul2(gq[c], a.rate2), neg), v.x, v.y);
					reinterpret_cast<float2*>(a.warp[c])[st] = v;
				}
			}
			f32x2 sq = mul2(gq[0], gq[0]);
			sq = add2(sq, mul2(gq[1], gq[1]), one);
			sq = add2(sq, mul2(gq[2], gq[2]), one);
			float sq_lo, sq_hi;
			unpack2(sq, sq_lo, sq_hi);
			best = fmaxf(best, fmaxf(sq_lo, sq_hi));  // NaNs are ignored, like `if (sq > best)` (statistics.tpp:65)
		}
		st += ZP;
	};

	f32x2 va[3], vb[3];
	request(va);
	request(vb);
	int r = r_first;
	// priming: rows whose consumption completes no output row of this block yet (at most 2R of them)
#pragma unroll 1
	for (; r < ys + R; r++) {
		chain(va);
#pragma unroll
		for (int c = 0; c < 3; c++) va[c] = vb[c];
		request(vb);
	}
	// steady state, two rows per trip: consuming row r completes output row r - R. First the trips whose two requests
	// (rows r + 2 and r + 3) lie inside the volume, then the last ones, whose requests may fall behind the last row.
#pragma unroll 1
	for (; r + 3 < r_load_end; r += 2) {
		chain(va);
		request_inside(va);
		emit(std::integral_constant<int, 0>());
		chain(vb);
		request_inside(vb);
		emit(std::integral_constant<int, 1>());
	}
#pragma unroll 1
	for (; r + 1 < r_stop; r += 2) {
		chain(va);
		request(va);
		emit(std::integral_constant<int, 0>());
		chain(vb);
		request(vb);
		emit(std::integral_constant<int, 1>());
	}
	if (r < r_stop) {
		chain(va);
		emit(std::integral_constant<int, 0>());
	}
	if (slots != nullptr) block_atomic_max(best, slots + a.iteration);
	__syncthreads();  // the next segment starts in row buffer 0 again
	}
}

// LSF_YMARCH3=0 keeps the fourth-generation filter kernel (A/B parity tests)
inline bool ymarch3_enabled() {
	const char* e = getenv("LSF_YMARCH3");
	return !(e && e[0] == '0');
}

inline bool ymarch3_supported(const Grid3& g, const float* h, const float* filtered, const float* warp) {
	return ymarch3_enabled() && g.Z >= 128 && ymarch2_supported(g, h, filtered, warp) && g.N * 3 < (1ll << 31);
}

template<int R>
void launch_ymarch3(const Taps& taps, const HierIterArgs& a, const float* h, float* filtered, float* warp, int y_chunk,
		cudaStream_t stream, int x_begin = 0, int planes = -1) {
	const Grid3& g = a.g;
	YMarch3Args f;
	for (int c = 0; c < 3; c++) {
		f.in[c] = h + c * g.N;
		f.out[c] = filtered ? filtered + c * g.N : nullptr;
		f.warp[c] = warp ? warp + c * g.N : nullptr;
	}
	f.X = g.X;
	f.Y = g.Y;
	f.Z = g.Z;
	for (int q = 0; q < 7; q++) f.k2[q] = dup2(q < 2 * R + 1 ? taps.k[q] : 0.0f);
	f.one2 = dup2(1.0f);
	f.neg2 = dup2(-1.0f);
	f.rate2 = dup2(a.rate);
	f.threshold = a.threshold;
	f.max_sq_bits = a.max_sq_bits;
	f.iteration = a.iteration;
	f.check_convergence = a.check_convergence;
	f.x_begin = x_begin;
	f.batch_X = a.batch_X;
	f.batch_slot_stride = a.batch_slot_stride;
	const int tile_z = g.Z < 256 ? 128 : 256;
	const int tiles = div_up(g.Z, tile_z);
	const int plane_count = planes < 0 ? g.X : planes;
	// resident blocks per SM (register-bound): 6 of 128 threads, 12 of 64; the variants that update the warp need 96
	// registers: 5 and 10 (768 blocks on 740 slots would run a second, almost empty wave: measured 0.18 against 0.13 ms)
	const int per_sm = (warp != nullptr ? 5 : 6) * (256 / tile_z);
	if (warp != nullptr || tile_z != 256) y_chunk = marching_chunk(g.Y, tiles * plane_count, 2 * R, per_sm);
	f.y_chunk = y_chunk;
	f.planes = plane_count;
	dim3 grid(tiles, plane_count, div_up(g.Y, y_chunk));
	// balanced mode: when the (plane, y-chunk) grid would leave part of the SMs one resident block short, the rows of all
	// planes are split evenly over exactly one wave of blocks instead (LSF_YM3_BALANCE=0: always y-chunks)
	{
		static int sm_count = 0;
		if (sm_count == 0) {
			int device = 0;
			cudaGetDevice(&device);
			if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sm_count <= 0)
				sm_count = 148;
		}
		const char* e = getenv("LSF_YM3_BALANCE");
		const char* forced = getenv("LSF_YM3_BLOCKS");  // experiment: blocks per SM of the balanced grid
		const long long slots_total = (long long) sm_count * (forced && atoi(forced) > 0 ? atoi(forced) : per_sm);
		const long long blocks = (long long) grid.x * grid.y * grid.z;
		const long long share = (long long) plane_count * g.Y * tiles / slots_total;  // rows per block
		if (!(e && e[0] == '0') && (forced || blocks <= slots_total) && blocks % sm_count != 0 && share >= 8 * R && share <= g.Y
				&& slots_total % tiles == 0) {
			f.y_chunk = 0;
			grid = dim3(tiles, (unsigned) (slots_total / tiles), 1);
		}
	}
	const int threads = tile_z / 2 + (tiles > 1 ? 32 : 0);
	const bool sym = taps_are_symmetric(taps);  // LSF_SYM=0: full chain (A/B)
#define LSF_YM3(SYM, OUT, WARP)                                                                          \
	do {                                                                                                 \
		if (tile_z == 256) launch_dependent(k_sobolev_ymarch3<R, SYM, OUT, WARP, 256>, grid, dim3(threads), 0, stream, f); \
		else launch_dependent(k_sobolev_ymarch3<R, SYM, OUT, WARP, 128>, grid, dim3(threads), 0, stream, f);             \
	} while (0)
	if (filtered && warp) {
		if (sym) LSF_YM3(true, true, true);
		else LSF_YM3(false, true, true);
	} else if (filtered) {
		if (sym) LSF_YM3(true, true, false);
		else LSF_YM3(false, true, false);
	} else {
		if (sym) LSF_YM3(true, false, true);
		else LSF_YM3(false, false, true);
	}
#undef LSF_YM3
}

template<int R>
void launch_ymarch_auto(const Taps& taps, const HierIterArgs& a, const float* h, float* filtered, float* warp, int y_chunk,
		cudaStream_t stream, int x_begin, int planes) {
	if (ymarch3_supported(a.g, h, filtered ? filtered : h, warp ? warp : h))
		launch_ymarch3<R>(taps, a, h, filtered, warp, y_chunk, stream, x_begin, planes);
	else launch_ymarch2<R>(taps, a, h, filtered, warp, y_chunk, stream, x_begin, planes);
}

#endif  // __CUDACC__

}  // namespace lsf
