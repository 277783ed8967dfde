// slavcheva_fast.cuh -- second generation of the 3D SobolevFusion / KillingFusion Sobolev filter (sm_100a).
//
// First generation (slavcheva.cuh: k_slav_filter_axis x 3) moves 3 x 24 B per voxel and re-reads every tap through the
// cache: 3 x 194 us per 256^3 iteration, 1.8 TB/s (profiles/r1_killing_v4.md). Here the three passes of
// convolve_with_kernel_preserve_zeros (reference cpp/src/math/convolution.cpp:23-67,69-145, C++ zero rule: a voxel whose
// pass-input vector is exactly zero yields the zero vector) run in two marching kernels on voxel pairs with packed
// f32x2 chains, the organisation of kernels3d_pair.cuh:
//   k_slav_xmarch<R>   axis-0 pass: a thread owns a z pair of one row and marches along axis 0 (pure streaming, the
//                      partial sums of the K outputs in flight live in registers);
//   k_slav_ymarch2<R>  axis-1 and axis-2 passes: marching along axis 1, axis-2 taps from a double-buffered shared row
//                      kept twice (as is / shifted by one column) so that every operand pair is one aligned LDS.64.
// Arithmetic per output: sum over taps q ascending of in[i - R + q] * k[q], zeros outside the field -- the order of
// k_slav_filter_axis; results are bit-identical (tests/test_gpu_parity_slavcheva.py).
#pragma once

#include "kernels3d_pair.cuh"
#include "slavcheva.cuh"

#include <cstring>

namespace lsf {

struct SlavMarchArgs {
	const float* in;   // planes
	float* out;        // planes
	int X, Y, Z;
	unsigned long long k2[7];  // taps duplicated into both lanes: k2[q] multiplies in[i - R + q]
	unsigned long long one2;
	const int* status;
	int iteration;
	int chunk;         // outputs per block along the marching axis
	int tile_z;        // k_slav_ymarch2: output columns per block (even)
};

#ifdef __CUDACC__

// zero flags of a voxel pair: bit 0 = low voxel's vector is exactly zero, bit 1 = high voxel's
__device__ __forceinline__ unsigned pair_zero_flags(f32x2 v0, f32x2 v1, f32x2 v2) {
	float a0, a1, b0, b1, c0, c1;
	unpack2(v0, a0, a1);
	unpack2(v1, b0, b1);
	unpack2(v2, c0, c1);
	return ((a0 == 0.0f && b0 == 0.0f && c0 == 0.0f) ? 1u : 0u) | ((a1 == 0.0f && b1 == 0.0f && c1 == 0.0f) ? 2u : 0u);
}
__device__ __forceinline__ f32x2 pair_clear(f32x2 v, unsigned flags) {
	float lo, hi;
	unpack2(v, lo, hi);
	return pack2((flags & 1u) ? 0.0f : lo, (flags & 2u) ? 0.0f : hi);
}

template<int R>
static __global__ void __launch_bounds__(256) k_slav_xmarch(const __grid_constant__ SlavMarchArgs a) {
	constexpr int K = 2 * R + 1;
	if (a.status[a.iteration]) return;
	const int pairs_per_plane = a.Y * a.Z / 2;
	const int pair = blockIdx.x * blockDim.x + threadIdx.x;
	if (pair >= pairs_per_plane) return;
	const int YZ = a.Y * a.Z;
	const long long N = (long long) a.X * YZ;
	const int xs = blockIdx.y * a.chunk;
	const int xe = min(a.X, xs + a.chunk);
	const int x_first = max(xs - R, 0), x_stop = xe + R;
	const f32x2 one = a.one2;
	f32x2 acc[3][K];
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int q = 0; q < K; q++) acc[c][q] = 0ull;
	unsigned flags = 0;  // 2 bits per plane, newest in the low bits
	long long at = (long long) x_first * YZ + 2 * pair;
	f32x2 next[3] = { 0ull, 0ull, 0ull };
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
		flags = (flags << 2) | pair_zero_flags(v0, v1, v2);  // planes >= X: zero vectors (never written)
		// plane x is tap q of output plane x + R - q
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
			const unsigned zero = (flags >> (2 * R)) & 3u;  // flags of plane x - R
			const long long o = at - (long long) R * YZ;
#pragma unroll
			for (int c = 0; c < 3; c++) {
				float2 v;
				unpack2(pair_clear(acc[c][K - 1], zero), v.x, v.y);
				*reinterpret_cast<float2*>(a.out + c * N + o) = v;
			}
		}
	}
}

// REWARP: the masked re-warp of the live field (k_slav_resample: reference warp_2d_advanced, field_warping.cpp:64-136,
// + maximum warp length) runs in the epilogue on the voxel pair's freshly filtered update vectors, so the filtered
// field is neither written nor read back (a.out may be nullptr).
template<int R, bool REWARP>
static __global__ void __launch_bounds__(288) k_slav_ymarch2(const __grid_constant__ SlavMarchArgs a,
		const __grid_constant__ SlavResampleArgs ra) {
	constexpr int K = 2 * R + 1;
	constexpr int H = 4;  // halo columns kept either side of the tile (>= R, even)
	if (a.status[a.iteration]) return;
	extern __shared__ __align__(16) float slav_rows[];  // [2 buffers][3 components][A: W | B: W]
	const int Y = a.Y, Z = a.Z;
	const int NT = a.tile_z / 2;
	const int W = a.tile_z + 2 * H;
	const int tid = threadIdx.x;
	const int z0 = blockIdx.x * a.tile_z;
	const int x = blockIdx.y;
	const int ys = blockIdx.z * a.chunk;
	const int ye = min(Y, ys + a.chunk);
	const bool owner = tid < NT;
	const int j = tid - NT;
	const int il = owner ? H + 2 * tid : (j < H / 2 ? 2 * j : a.tile_z + H + 2 * (j - H / 2));
	const int z = z0 - H + il;
	const bool active = (owner || j < H) && z >= 0 && z < Z;
	const bool writes = owner && active;
	for (int i = tid; i < 12 * W; i += blockDim.x) slav_rows[i] = 0.0f;  // columns outside the field stay zero
	__syncthreads();
	const uint32_t rows = smem_addr(slav_rows);
	const f32x2 one = a.one2;
	f32x2 acc[3][K];
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int q = 0; q < K; q++) acc[c][q] = 0ull;
	unsigned flags = 0;
	const int r_first = max(ys - R, 0), r_stop = ye + R, r_load_end = min(r_stop, Y);
	const long long N = (long long) a.X * Y * Z;
	long long at = ((long long) x * Y + r_first) * Z + z;
	f32x2 next[3] = { 0ull, 0ull, 0ull };
	if (active) {
#pragma unroll
		for (int c = 0; c < 3; c++) {
			const float2 v = __ldg(reinterpret_cast<const float2*>(a.in + c * N + at));
			next[c] = pack2(v.x, v.y);
		}
	}
	int buffer = 0;
	float sq_report = 0.0f;
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
		flags = (flags << 2) | pair_zero_flags(v0, v1, v2);
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
		const long long o = at - (long long) R * Z;  // voxel (x, r - R, z)
		// axis-1 result of this pair, with the axis-1 zero rule; it is the input (and the zero test) of the axis-2 pass
		const unsigned zero1 = (flags >> (2 * R)) & 3u;
		f32x2 mid[3];
#pragma unroll
		for (int c = 0; c < 3; c++) mid[c] = pair_clear(acc[c][K - 1], zero1);
		if (active) {
#pragma unroll
			for (int c = 0; c < 3; c++) {
				const uint32_t pa = row + (c * 2 * W + il) * 4;
				float lo, hi;
				unpack2(mid[c], lo, hi);
				sts_f32x2(pa, mid[c]);
				if (il > 0) sts_f32(pa + (W - 1) * 4, lo);  // B[il - 1] = A[il]
				sts_f32(pa + W * 4, hi);                      // B[il] = A[il + 1]
			}
		}
		__syncthreads();
		if (writes) {
			const unsigned zero2 = pair_zero_flags(mid[0], mid[1], mid[2]);
			float update_lo[3], update_hi[3];
#pragma unroll
			for (int c = 0; c < 3; c++) {
				const uint32_t pa = row + (c * 2 * W + il - R) * 4;
				f32x2 sum = 0ull;
#pragma unroll
				for (int q = 0; q < K; q++) {
					const f32x2 v = ((q - R) % 2 == 0) ? lds_f32x2(pa + q * 4) : lds_f32x2(pa + (W + q - 1) * 4);
					sum = q == 0 ? mul2(v, a.k2[0]) : add2(sum, mul2(v, a.k2[q]), one);
				}
				unpack2(pair_clear(sum, zero2), update_lo[c], update_hi[c]);
				if (a.out != nullptr) *reinterpret_cast<float2*>(a.out + c * N + o) = make_float2(update_lo[c], update_hi[c]);
			}
			if (REWARP) {
				const float2 live2 = __ldg(reinterpret_cast<const float2*>(ra.live + o));
				const float2 canonical2 = ra.band_union_only ? __ldg(reinterpret_cast<const float2*>(ra.canonical + o))
						: make_float2(0.f, 0.f);
				float new_lo, new_hi, w_lo[3], w_hi[3];
				slav_resample_voxel<3>(ra, (int) o, update_lo, live2.x, canonical2.x, new_lo, w_lo, sq_report);
				slav_resample_voxel<3>(ra, (int) o + 1, update_hi, live2.y, canonical2.y, new_hi, w_hi, sq_report);
				*reinterpret_cast<float2*>(ra.new_live + o) = make_float2(new_lo, new_hi);
#pragma unroll
				for (int c = 0; c < 3; c++) *reinterpret_cast<float2*>(ra.warp + c * N + o) = make_float2(w_lo[c], w_hi[c]);
			}
		}
		buffer ^= 1;
	}
	if (REWARP && ra.max_sq_bits != nullptr) block_atomic_max(sq_report, ra.max_sq_bits);
}

inline bool slav_fast_filter_supported(const SlavGeom& g, const Taps& taps, const float* a, const float* b) {
	auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; };
	return g.nd == 3 && taps.radius >= 1 && taps.radius <= 3 && g.n[2] % 2 == 0 && g.N * 3 < (1ll << 31) && aligned(a)
			&& aligned(b);
}

// in -> scratch (axis 0) -> out (axes 1, 2); returns the number of launches
// `rewarp` (optional): run the masked re-warp in the second kernel's epilogue instead of writing the filtered field
template<int R>
int launch_slav_fast_filter(const SlavGeom& g, const Taps& taps, const float* in, float* scratch, float* out,
		const int* status, int iteration, cudaStream_t stream, const SlavResampleArgs* rewarp = nullptr) {
	SlavMarchArgs f;
	f.X = g.n[0];
	f.Y = g.n[1];
	f.Z = g.n[2];
	for (int q = 0; q < 7; q++) f.k2[q] = dup2(q < 2 * R + 1 ? taps.k[q] : 0.0f);
	f.one2 = dup2(1.0f);
	f.status = status;
	f.iteration = iteration;
	// axis 0: one thread per z pair of a plane, chunks along x sized to fill the GPU a few times over
	f.in = in;
	f.out = scratch;
	const int pairs = f.Y * f.Z / 2;
	const int plane_blocks = (int) div_up(pairs, 256);
	f.chunk = marching_chunk(f.X, plane_blocks, 2 * R, 3);
	f.tile_z = 0;
	k_slav_xmarch<R> <<<counted(dim3(plane_blocks, (unsigned) div_up(f.X, f.chunk))), 256, 0, stream>>>(f);
	// axes 1 and 2
	f.in = scratch;
	f.out = out;
	f.tile_z = std::min(512, (int) div_up(f.Z, 64) * 64);
	const int tiles = (int) div_up(f.Z, f.tile_z);
	f.chunk = marching_chunk(f.Y, tiles * f.X, 2 * R, 6);
	const int threads = f.tile_z / 2 + (tiles > 1 ? 32 : 0);
	const size_t shared = (size_t) 12 * (f.tile_z + 8) * sizeof(float);
	const dim3 grid(tiles, f.X, (unsigned) div_up(f.Y, f.chunk));
	if (rewarp != nullptr) {
		f.out = nullptr;
		k_slav_ymarch2<R, true> <<<counted(grid), threads, shared, stream>>>(f, *rewarp);
	} else {
		SlavResampleArgs none;
		std::memset(&none, 0, sizeof(none));
		k_slav_ymarch2<R, false> <<<counted(grid), threads, shared, stream>>>(f, none);
	}
	return 2;
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------- narrow-band sparse iteration
// The C++-semantics 3D iteration touches, in effect, only the narrow band: a voxel whose live and canonical values are
// both truncated gets a zero update (data_term.cpp:72-83, smoothing_term.cpp:43-108 skip it), the zero rule of
// convolve_with_kernel_preserve_zeros keeps its filtered update at zero (convolution.cpp:23-67), and the re-warp leaves
// its live value alone (field_warping.cpp:88-91) -- and since such a voxel never changes again, the band only shrinks.
// The sparse iteration keeps these invariants in the buffers instead of re-establishing them every iteration:
//   * outside the band the three update fields and the warp are zero and the two live buffers hold the same value
//     (set up once per optimize(); a voxel that leaves the band is patched by k_slav_band_leave);
//   * k_slav_band_scan classifies the voxels of the two scalar fields (blocks of 1024 voxels that hold no band voxel are
//     flagged and never looked at again) and writes the band list in memory order; k_slav_band_terms evaluates the
//     gradient terms at the listed voxels in full warps; the list can be re-used for several iterations (LSF_SLAV_RESCAN; voxels
//     that have left the band meanwhile are recognised and handled like the dense path handles them);
//   * the three filter passes, the re-warp and the maximum warp length run over that list.
// Per-voxel arithmetic is that of the dense kernels (same device functions); taps that fall outside the band read the
// zeros the dense path would have computed there.

struct SlavBandArgs {
	SlavGeom g;
	int* list;          // band voxels of this iteration (blocks in arbitrary order, memory order inside a block)
	int* positions;     // their coordinates, 10 bits per axis (x | y << 10 | z << 20); every dimension is <= 1024
	int* count;         // their number (slot of this iteration, zero before the gradient kernel)
	unsigned char* dead;  // per block of 1024 voxels: no band voxel left (the band only shrinks: never looked at again)
	int* leave_list;    // voxels that left the band in this iteration's re-warp
	int* leave_count;
	const int* status;
	int iteration;
};

#ifdef __CUDACC__

static __global__ void __launch_bounds__(256) k_slav_band_scan(SlavGradientArgs a, SlavBandArgs b) {
	if (a.status[a.iteration] || b.dead[blockIdx.x]) return;
	__shared__ unsigned short band_local[1024];
	__shared__ int warp_totals[8];
	__shared__ int band_count, band_base;
	const long long block_base = (long long) blockIdx.x * 1024;
	const long long first = block_base + threadIdx.x * 4;
	unsigned in_band = 0;
	if (first < a.g.N) {
		const int base = (int) first;
		const float4 live4 = __ldg(reinterpret_cast<const float4*>(a.live + base));
		const float4 canonical4 = __ldg(reinterpret_cast<const float4*>(a.canonical + base));
		const float live_v[4] = { live4.x, live4.y, live4.z, live4.w };
		const float canonical_v[4] = { canonical4.x, canonical4.y, canonical4.z, canonical4.w };
#pragma unroll
		for (int v = 0; v < 4; v++)
			if (!(slav_truncated(live_v[v]) && slav_truncated(canonical_v[v]))) in_band |= 1u << v;
	}
	// ordered compaction (exclusive scan over the block): the list keeps the voxels in memory order, so that
	// neighbouring threads of the list kernels touch neighbouring addresses
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int mine = __popc(in_band);
	int inclusive = mine;
#pragma unroll
	for (int offset = 1; offset < 32; offset <<= 1) {
		const int other = __shfl_up_sync(0xffffffffu, inclusive, offset);
		if (lane >= offset) inclusive += other;
	}
	if (lane == 31) warp_totals[warp] = inclusive;
	__syncthreads();
	int before = inclusive - mine;
	for (int w = 0; w < warp; w++) before += warp_totals[w];
	if (threadIdx.x == 255) band_count = before + mine;
	{
		int k = 0;
#pragma unroll
		for (int v = 0; v < 4; v++)
			if (in_band & (1u << v)) band_local[before + k++] = (unsigned short) (threadIdx.x * 4 + v);
	}
	__syncthreads();
	const int count = band_count;
	if (count == 0) {
		if (threadIdx.x == 0) b.dead[blockIdx.x] = 1;
		return;
	}
	if (threadIdx.x == 0) band_base = atomicAdd(b.count, count);
	__syncthreads();
	for (int j = threadIdx.x; j < count; j += 256) {
		const int idx = (int) block_base + band_local[j];
		int q[3];
		slav_coords<3>(a.g, idx, q);
		b.list[band_base + j] = idx;
		b.positions[band_base + j] = q[0] | (q[1] << 10) | (q[2] << 20);
	}
}

// the gradient terms at one listed band voxel (packed = its coordinates, 10 bits per axis)
__device__ __forceinline__ void slav_band_terms_voxel(const SlavGradientArgs& a, int idx, int packed) {
	const SlavParams& p = a.p;
	const bool killing = p.smoothing_term_method == LSF_SMOOTHING_KILLING;
	const int q[3] = { packed & 1023, (packed >> 10) & 1023, (packed >> 20) & 1023 };
	const float live_value = __ldg(a.live + idx);
	if (slav_truncated(live_value) && slav_truncated(__ldg(a.canonical + idx))) {
		// the list is re-used for several iterations: this voxel has left the band since the last scan
#pragma unroll
		for (int c = 0; c < 3; c++) a.out[c * a.g.N + idx] = (0.0f + 0.0f * p.smoothing_weight) * -p.rate;
		return;
	}
	float data[3], smooth[3], ls[3];
	const bool ls_here = p.level_set && !slav_truncated(live_value);
	const bool interior = q[0] >= 1 && q[0] < a.g.n[0] - 1 && q[1] >= 1 && q[1] < a.g.n[1] - 1 && q[2] >= 1
			&& q[2] < a.g.n[2] - 1;
	if (interior) {
		slav_data_term<3, true>(a, idx, q, data);
		if (killing) slav_killing<3, true>(a, idx, q, smooth);
		else slav_tikhonov_cpp<3, true>(a, idx, q, smooth);
		if (ls_here) slav_level_set<3, true>(a, idx, q, ls);
	} else {
		slav_data_term<3>(a, idx, q, data);
		if (killing) slav_killing<3>(a, idx, q, smooth);
		else slav_tikhonov_cpp<3>(a, idx, q, smooth);
		if (ls_here) slav_level_set<3>(a, idx, q, ls);
	}
#pragma unroll
	for (int c = 0; c < 3; c++) {
		float total = data[c] * p.data_weight;
		if (ls_here) total = total + ls[c] * p.level_set_weight;
		total = total + smooth[c] * p.smoothing_weight;
		a.out[c * a.g.N + idx] = total * -p.rate;  // reference sobolev_optimizer2d.cpp:131-132
	}
}

// the gradient terms at the listed band voxels (full warps: the list is dense)
static __global__ void __launch_bounds__(256) k_slav_band_terms(SlavGradientArgs a, SlavBandArgs b) {
	if (a.status[a.iteration]) return;
	const int count = *b.count;
	for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < count; j += gridDim.x * blockDim.x)
		slav_band_terms_voxel(a, b.list[j], b.positions[j]);
}

// one pass of convolve_with_kernel_preserve_zeros (C++ zero rule) at one band voxel (i = its position along the pass
// axis); `in` is zero outside the band
template<int R>
__device__ __forceinline__ void slav_band_filter_voxel(const SlavFilterArgs& a, int idx, int i, float (&acc)[3]) {
	const int n = a.g.n[a.axis], s = a.g.stride[a.axis];
	const int N = (int) a.g.N;
	bool all_zero = true;
#pragma unroll
	for (int c = 0; c < 3; c++) all_zero = all_zero && __ldg(a.in + c * N + idx) == 0.0f;
	acc[0] = acc[1] = acc[2] = 0.0f;
	if (all_zero) return;
	if (i >= R && i < n - R) {  // every tap inside the field
#pragma unroll
		for (int c = 0; c < 3; c++) {
			const float* line = a.in + c * N + idx;
#pragma unroll
			for (int t = 0; t < 2 * R + 1; t++) acc[c] += __ldg(line + (t - R) * s) * a.k[t];
		}
	} else {
#pragma unroll
		for (int c = 0; c < 3; c++) {
			const float* line = a.in + c * N + idx;
#pragma unroll
			for (int t = 0; t < 2 * R + 1; t++) {
				const int src = i - R + t;
				const float value = (src >= 0 && src < n) ? __ldg(line + (t - R) * s) : 0.0f;
				acc[c] += value * a.k[t];
			}
		}
	}
}

template<int R>
static __global__ void __launch_bounds__(256) k_slav_band_filter_axis(SlavFilterArgs a, SlavBandArgs b) {
	if (a.status[a.iteration]) return;
	const int count = *b.count;
	const int N = (int) a.g.N;
	for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < count; j += gridDim.x * blockDim.x) {
		const int idx = b.list[j];
		float acc[3];
		slav_band_filter_voxel<R>(a, idx, (b.positions[j] >> (10 * a.axis)) & 1023, acc);
#pragma unroll
		for (int c = 0; c < 3; c++) a.out[c * N + idx] = acc[c];
	}
}

// re-warp of the band voxels (k_slav_resample's per-voxel function) + maximum warp length; voxels whose new value is
// truncated while the canonical one is too have left the band for good and are queued for k_slav_band_leave
static __global__ void __launch_bounds__(256) k_slav_band_resample(SlavResampleArgs a, SlavBandArgs b) {
	if (a.status != nullptr && a.status[a.iteration]) return;
	const int count = *b.count;
	const SlavGeom& g = a.g;
	float sq_report = 0.0f;
	for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < count; j += gridDim.x * blockDim.x) {
		const int idx = b.list[j];
		float update[3], w[3], new_value;
#pragma unroll
		for (int c = 0; c < 3; c++) update[c] = __ldg(a.update + c * g.N + idx);
		const float canonical_value = __ldg(a.canonical + idx);
		slav_resample_voxel<3>(a, idx, update, __ldg(a.live + idx), canonical_value, new_value, w, sq_report);
		a.new_live[idx] = new_value;
#pragma unroll
		for (int c = 0; c < 3; c++) a.warp[c * g.N + idx] = w[c];
		if (slav_truncated(new_value) && slav_truncated(canonical_value)) b.leave_list[atomicAdd(b.leave_count, 1)] = idx;
	}
	if (a.max_sq_bits != nullptr) block_atomic_max(sq_report, a.max_sq_bits);
}

// restores the invariants at the voxels that left the band: both live buffers equal, update fields zero
static __global__ void __launch_bounds__(256) k_slav_band_leave(SlavBandArgs b, const float* new_live, float* old_live,
		float* field_a, float* field_b, float* field_f) {
	if (b.status[b.iteration]) return;
	const int count = *b.leave_count;
	for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < count; j += gridDim.x * blockDim.x) {
		const int idx = b.leave_list[j];
		old_live[idx] = new_live[idx];
#pragma unroll
		for (int c = 0; c < 3; c++) {
			field_a[c * b.g.N + idx] = 0.0f;
			field_b[c * b.g.N + idx] = 0.0f;
			field_f[c * b.g.N + idx] = 0.0f;
		}
	}
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------- brick-ordered narrow band
// Fifth generation of the sparse iteration (default). ncu of the list kernels above (profiles/r1_killing_v4.md): every
// band voxel issues 21 - 60 dependent 4-byte loads that no neighbouring thread shares (the list is in memory order, the
// x / y neighbours of a voxel sit in other blocks on other SMs): 7x the pass input crosses from L2 to the SMs, L1 hit rate
// 14 %, 30 long-scoreboard stalls per issue. Here the band is organised in bricks of 8 x 8 x 32 voxels and every operand
// of a brick is staged in shared memory by the TMA unit:
//   * k_slav_brick_scan: one block per brick classifies its 2048 voxels (128-bit loads) and writes the brick's band
//     voxels (16-bit offsets inside the brick, rows of z-adjacent voxels in order) into the brick's own segment of the
//     list; bricks with band voxels are appended to the active list, bricks without are dead for good;
//   * the iteration kernels are persistent (one block per SM): a producer warp claims the next active brick from a
//     device-side cursor and fetches the brick's box of every input field with its halo (1 voxel for the term stencils,
//     R voxels along the pass axis for a filter pass; cp.async.bulk.tensor.4d, zero fill outside the volume = the
//     filter's zero padding), the canonical box and the brick's list segment (cp.async.bulk) into a ring of stages; the
//     consumer warps wait on a stage's `full` mbarrier, work through the list -- every stencil / filter tap is an LDS with
//     a compile-time offset from one base address, no global load is left -- and release the stage on its `empty`
//     mbarrier warp by warp (no block-wide barrier). The boxes of the next bricks are in flight while a brick is computed;
//   * the axis-2 pass and the masked re-warp run in one kernel (the filtered update of a voxel is all its re-warp
//     needs), the patch-up of the voxels that left the band and the termination test in another: 5 launches per
//     iteration instead of 7.
// Voxels on the faces of the volume (one-sided stencils) take the global path of the list kernels.
// Per-voxel arithmetic: the same device functions as the list kernels and the dense kernels, in the same order.
constexpr int SLAV_BRICK_X = 8, SLAV_BRICK_Y = 8, SLAV_BRICK_Z = 32;
constexpr int SLAV_BRICK_VOXELS = SLAV_BRICK_X * SLAV_BRICK_Y * SLAV_BRICK_Z;
constexpr int SLAV_STAGE_Z = SLAV_BRICK_Z + 8;  // staged rows start 4 voxels before the brick (16-byte aligned rows)

struct SlavBrickArgs {
	unsigned short* list;  // [bricks][2048]: band voxels of the brick as offsets inside it ((x * 8 + y) * 32 + z)
	int* brick_count;      // listed voxels per brick
	unsigned char* dead;   // brick holds no band voxel (the band only shrinks: never scanned again)
	int bricks_y, bricks_z;
	int* leave_list;       // voxels that left the band in this iteration's re-warp
	int* leave_count;
	int* active;           // bricks with band voxels, appended by the scan (arbitrary order)
	int* active_count;
	int* cursor;           // next entry of `active` to claim (zero at launch; one cursor per launch)
	const int* status;
	int iteration;
};

struct SlavBrickMaps {
	CUtensorMap live[2];    // [buffer parity] box 40 x 10 x 10 (z, y, x)
	CUtensorMap canonical;  // box 32 x 8 x 8
	CUtensorMap warp;       // box 40 x 10 x 10 x 3
	CUtensorMap pass[3];    // input field of filter pass `axis`: brick + R-voxel halo along the axis (z rows of 40 for axis 2)
};

// box extents (x, y, z) of the staged input of filter pass AXIS
template<int R, int AXIS> struct SlavPassBox {
	static constexpr int X = SLAV_BRICK_X + (AXIS == 0 ? 2 * R : 0);
	static constexpr int Y = SLAV_BRICK_Y + (AXIS == 1 ? 2 * R : 0);
	static constexpr int Z = AXIS == 2 ? SLAV_STAGE_Z : SLAV_BRICK_Z;
	static constexpr int LO_X = AXIS == 0 ? R : 0, LO_Y = AXIS == 1 ? R : 0, LO_Z = AXIS == 2 ? 4 : 0;
	static constexpr int VOXELS = X * Y * Z;
	static constexpr int STRIDE = AXIS == 0 ? Y * Z : (AXIS == 1 ? Z : 1);
};

inline int make_brick_box_map(CUtensorMap* map, const float* base, int channels, long long channel_stride, const SlavGeom& g,
		int box_x, int box_y, int box_z) {
	EncodeTiledFn encode = encode_tiled_fn();
	LSF_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
	const cuuint64_t dims[4] = { (cuuint64_t) g.n[2], (cuuint64_t) g.n[1], (cuuint64_t) g.n[0], (cuuint64_t) channels };
	const cuuint64_t strides[3] = { (cuuint64_t) g.n[2] * 4, (cuuint64_t) g.n[1] * g.n[2] * 4, (cuuint64_t) channel_stride * 4 };
	const cuuint32_t box[4] = { (cuuint32_t) box_z, (cuuint32_t) box_y, (cuuint32_t) box_x, (cuuint32_t) channels };
	const cuuint32_t element_strides[4] = { 1, 1, 1, 1 };
	const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box,
			element_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
			CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	LSF_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d (brick box %d x %d x %d)", (int) r, box_x, box_y,
			box_z);
	return LSF_OK;
}

#ifdef __CUDACC__

// exclusive prefix sum of `mine` over the 256 threads of the block; *total = sum (needs two barriers)
__device__ __forceinline__ int slav_block_exclusive_scan(int mine, int* warp_totals, int* total) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int inclusive = mine;
#pragma unroll
	for (int offset = 1; offset < 32; offset <<= 1) {
		const int other = __shfl_up_sync(0xffffffffu, inclusive, offset);
		if (lane >= offset) inclusive += other;
	}
	__syncthreads();  // warp_totals may still be read from a previous call
	if (lane == 31) warp_totals[warp] = inclusive;
	__syncthreads();
	int before = inclusive - mine, sum = 0;
	for (int w = 0; w < 8; w++) {
		if (w < warp) before += warp_totals[w];
		sum += warp_totals[w];
	}
	*total = sum;
	return before;
}

__device__ __forceinline__ void slav_brick_origin(const SlavBrickArgs& b, int brick, int& x0, int& y0, int& z0) {
	z0 = (brick % b.bricks_z) * SLAV_BRICK_Z;
	y0 = ((brick / b.bricks_z) % b.bricks_y) * SLAV_BRICK_Y;
	x0 = (brick / (b.bricks_z * b.bricks_y)) * SLAV_BRICK_X;
}

static __global__ void __launch_bounds__(256) k_slav_brick_scan(SlavGradientArgs a, SlavBrickArgs b) {
	const int brick = blockIdx.x;
	if (a.status[a.iteration] || b.dead[brick]) return;
	__shared__ int warp_totals[8];
	int x0, y0, z0;
	slav_brick_origin(b, brick, x0, y0, z0);
	const int lz = (threadIdx.x & 7) * 4;
	int listed = 0;
	const int base = brick * SLAV_BRICK_VOXELS;
#pragma unroll
	for (int half = 0; half < 2; half++) {
		const int row = (threadIdx.x >> 3) + 32 * half;  // row of the brick: x-major
		const int x = x0 + (row >> 3), y = y0 + (row & 7), z = z0 + lz;
		unsigned in_band = 0;
		if (x < a.g.n[0] && y < a.g.n[1] && z < a.g.n[2]) {
			const int first = (x * a.g.n[1] + y) * a.g.n[2] + z;
			const float4 live4 = __ldg(reinterpret_cast<const float4*>(a.live + first));
			const float4 canonical4 = __ldg(reinterpret_cast<const float4*>(a.canonical + first));
			const float live_v[4] = { live4.x, live4.y, live4.z, live4.w };
			const float canonical_v[4] = { canonical4.x, canonical4.y, canonical4.z, canonical4.w };
#pragma unroll
			for (int v = 0; v < 4; v++)
				if (!(slav_truncated(live_v[v]) && slav_truncated(canonical_v[v]))) in_band |= 1u << v;
		}
		int total;
		int at = base + listed + slav_block_exclusive_scan(__popc(in_band), warp_totals, &total);
#pragma unroll
		for (int v = 0; v < 4; v++)
			if (in_band & (1u << v)) b.list[at++] = (unsigned short) (row * SLAV_BRICK_Z + lz + v);
		listed += total;
	}
	if (threadIdx.x == 0) {
		b.brick_count[brick] = listed;
		if (listed == 0) b.dead[brick] = 1;
		else b.active[atomicAdd(b.active_count, 1)] = brick;
	}
}

// restores the invariants at the voxels that left the band (k_slav_band_leave) and evaluates the termination test
// (k_slav_decide) in one launch
static __global__ void __launch_bounds__(256) k_slav_brick_leave_decide(SlavBrickArgs b, long long N, const float* new_live,
		float* old_live, float* field_a, float* field_b, float* field_f, SlavParams p, const unsigned* max_sq_bits,
		int* status, int max_iterations) {
	if (status[b.iteration]) {
		if (blockIdx.x == 0 && threadIdx.x == 0) status[b.iteration + 1] = 1;
		return;
	}
	const int count = *b.leave_count;
	for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < count; j += gridDim.x * blockDim.x) {
		const int idx = b.leave_list[j];
		old_live[idx] = new_live[idx];
#pragma unroll
		for (int c = 0; c < 3; c++) {
			field_a[c * N + idx] = 0.0f;
			field_b[c * N + idx] = 0.0f;
			field_f[c * N + idx] = 0.0f;
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		const float max_warp = sqrtf(__uint_as_float(max_sq_bits[b.iteration]));
		status[b.iteration + 1] = slav_finished(p, b.iteration + 1, max_iterations, max_warp) ? 1 : 0;
	}
}

// plain bulk copy global -> shared (bytes: multiple of 16, both addresses 16-byte aligned), completion on `bar`
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

// ring of stages shared by the producer warp and the consumer warps of a persistent block
template<int STAGES>
struct SlavPipe {
	uint64_t full[STAGES];   // the stage's boxes have landed (or: no brick left)
	uint64_t empty[STAGES];  // every consumer warp is done with the stage
	int brick[STAGES];       // brick staged there; -1 = the list is exhausted
	int count[STAGES];       // its listed voxels
};

template<int STAGES>
__device__ __forceinline__ void slav_pipe_init(SlavPipe<STAGES>& pipe, int consumer_warps) {
	if (threadIdx.x == 0) {
		for (int s = 0; s < STAGES; s++) {
			mbar_init(&pipe.full[s], 1);
			mbar_init(&pipe.empty[s], consumer_warps);
		}
		mbar_fence_init();
	}
	__syncthreads();
}

// producer: waits until stage n % STAGES is free and claims the next active brick; returns it (-1 at the end of the list:
// the stage is then completed so that the consumers see the end marker). The brick's list segment is requested here.
template<int STAGES>
__device__ __forceinline__ int slav_pipe_claim(SlavPipe<STAGES>& pipe, const SlavBrickArgs& b, int total, int n,
		unsigned short* stage_list, uint32_t box_bytes) {
	const int s = n % STAGES;
	if (n >= STAGES) mbar_wait(&pipe.empty[s], ((n / STAGES) - 1) & 1);
	const int k = atomicAdd(b.cursor, 1);
	const int brick = k < total ? b.active[k] : -1;
	pipe.brick[s] = brick;
	if (brick < 0) {
		mbar_arrive(&pipe.full[s]);
		return brick;
	}
	const int count = b.brick_count[brick];
	pipe.count[s] = count;
	const uint32_t list_bytes = ((uint32_t) count * 2u + 15u) & ~15u;
	mbar_expect_tx(&pipe.full[s], box_bytes + list_bytes);
	bulk_load(stage_list, b.list + (size_t) brick * SLAV_BRICK_VOXELS, list_bytes, &pipe.full[s]);
	return brick;
}

template<int STAGES>
__device__ __forceinline__ void slav_pipe_release(SlavPipe<STAGES>& pipe, int s) {
	__syncwarp();
	if ((threadIdx.x & 31) == 0) mbar_arrive(&pipe.empty[s]);
}

constexpr int SLAV_LIST_BYTES = SLAV_BRICK_VOXELS * 2;                                    // staged list segment
constexpr int SLAV_TERMS_TILE = (SLAV_BRICK_X + 2) * (SLAV_BRICK_Y + 2) * SLAV_STAGE_Z;  // voxels of one staged field
// stage of the terms kernel: live [10][10][40], warp [3][10][10][40], canonical [8][8][32], list
constexpr int SLAV_TERMS_BOX_BYTES = (4 * SLAV_TERMS_TILE + SLAV_BRICK_VOXELS) * 4;
constexpr int SLAV_TERMS_STAGE_BYTES = SLAV_TERMS_BOX_BYTES + SLAV_LIST_BYTES;
constexpr int SLAV_TERMS_STAGES = 3, SLAV_CONSUMERS = 768;

// gradient terms of one staged brick, `workers` threads
__device__ __forceinline__ void slav_terms_brick(const SlavGradientArgs& a, const SlavBrickArgs& b, const float* tile,
		const float* canonical_box, const unsigned short* list, int count, int brick, int worker, int workers) {
	int x0, y0, z0;
	slav_brick_origin(b, brick, x0, y0, z0);
	// the argument block of the staged fields: component stride and strides of the tile
	SlavGradientArgs sa = a;
	sa.live = tile;
	sa.warp = tile + SLAV_TERMS_TILE;
	sa.g.N = SLAV_TERMS_TILE;
	sa.g.stride[0] = (SLAV_BRICK_Y + 2) * SLAV_STAGE_Z;
	sa.g.stride[1] = SLAV_STAGE_Z;
	sa.g.stride[2] = 1;
	const SlavParams& p = a.p;
	const bool killing = p.smoothing_term_method == LSF_SMOOTHING_KILLING;
	const int N = (int) a.g.N;
	for (int j = worker; j < count; j += workers) {
		const int local = list[j];
		const int lx = local >> 8, ly = (local >> 5) & 7, lz = local & 31;
		const int q[3] = { x0 + lx, y0 + ly, z0 + lz };
		const int idx = (q[0] * a.g.n[1] + q[1]) * a.g.n[2] + q[2];
		const bool interior = q[0] >= 1 && q[0] < a.g.n[0] - 1 && q[1] >= 1 && q[1] < a.g.n[1] - 1 && q[2] >= 1
				&& q[2] < a.g.n[2] - 1;
		if (!interior) {
			slav_band_terms_voxel(a, idx, q[0] | (q[1] << 10) | (q[2] << 20));
			continue;
		}
		const int s[3] = { lx + 1, ly + 1, lz + 4 };
		const int at = (s[0] * (SLAV_BRICK_Y + 2) + s[1]) * SLAV_STAGE_Z + s[2];
		const float live_value = tile[at];
		const float canonical_value = canonical_box[local];
		if (slav_truncated(live_value) && slav_truncated(canonical_value)) {
			// the list is re-used for several iterations: this voxel has left the band since the last scan
#pragma unroll
			for (int c = 0; c < 3; c++) a.out[c * N + idx] = (0.0f + 0.0f * p.smoothing_weight) * -p.rate;
			continue;
		}
		float data[3], smooth[3], ls[3];
		const bool ls_here = p.level_set && !slav_truncated(live_value);
		slav_data_term_given<3, true, true>(sa, at, s, canonical_value, data);
		if (killing) slav_killing<3, true, true>(sa, at, s, smooth);
		else slav_tikhonov_cpp<3, true, true>(sa, at, s, smooth);
		if (ls_here) slav_level_set<3, true, true>(sa, at, s, ls);
#pragma unroll
		for (int c = 0; c < 3; c++) {
			float total = data[c] * p.data_weight;
			if (ls_here) total = total + ls[c] * p.level_set_weight;
			total = total + smooth[c] * p.smoothing_weight;
			a.out[c * N + idx] = total * -p.rate;  // reference sobolev_optimizer2d.cpp:131-132
		}
	}
}

static __global__ void __launch_bounds__(SLAV_CONSUMERS + 32, 1) k_slav_brick_terms_tma(SlavGradientArgs a, SlavBrickArgs b,
		const __grid_constant__ CUtensorMap live_map, const __grid_constant__ CUtensorMap warp_map,
		const __grid_constant__ CUtensorMap canonical_map) {
	if (a.status[a.iteration]) return;
	extern __shared__ __align__(128) unsigned char slav_smem[];
	__shared__ SlavPipe<SLAV_TERMS_STAGES> pipe;
	slav_pipe_init(pipe, SLAV_CONSUMERS / 32);
	if (threadIdx.x >= SLAV_CONSUMERS) {
		if (threadIdx.x == SLAV_CONSUMERS) {
			const int total = *b.active_count;
			for (int n = 0;; n++) {
				const int s = n % SLAV_TERMS_STAGES;
				unsigned char* stage = slav_smem + s * SLAV_TERMS_STAGE_BYTES;
				const int brick = slav_pipe_claim(pipe, b, total, n, reinterpret_cast<unsigned short*>(stage + SLAV_TERMS_BOX_BYTES),
						SLAV_TERMS_BOX_BYTES);
				if (brick < 0) break;
				int x0, y0, z0;
				slav_brick_origin(b, brick, x0, y0, z0);
				float* tile = reinterpret_cast<float*>(stage);
				tma_load_4d(tile, &live_map, z0 - 4, y0 - 1, x0 - 1, 0, &pipe.full[s]);
				tma_load_4d(tile + SLAV_TERMS_TILE, &warp_map, z0 - 4, y0 - 1, x0 - 1, 0, &pipe.full[s]);
				tma_load_4d(tile + 4 * SLAV_TERMS_TILE, &canonical_map, z0, y0, x0, 0, &pipe.full[s]);
			}
		}
		return;
	}
	for (int n = 0;; n++) {
		const int s = n % SLAV_TERMS_STAGES;
		mbar_wait(&pipe.full[s], (n / SLAV_TERMS_STAGES) & 1);
		const int brick = pipe.brick[s];
		if (brick < 0) break;
		const unsigned char* stage = slav_smem + s * SLAV_TERMS_STAGE_BYTES;
		const float* tile = reinterpret_cast<const float*>(stage);
		slav_terms_brick(a, b, tile, tile + 4 * SLAV_TERMS_TILE, reinterpret_cast<const unsigned short*>(stage + SLAV_TERMS_BOX_BYTES),
				pipe.count[s], brick, threadIdx.x, SLAV_CONSUMERS);
		slav_pipe_release(pipe, s);
	}
}

// one pass of convolve_with_kernel_preserve_zeros (C++ zero rule) at a band voxel from the staged box of the pass input
// (`at` = the voxel inside component 0 of the box); same tap order as slav_band_filter_voxel
template<int R, int AXIS>
__device__ __forceinline__ void slav_staged_filter_voxel(const float* box, int at, const float (&k)[LSF_MAX_KERNEL_SIZE],
		float (&acc)[3]) {
	typedef SlavPassBox<R, AXIS> Box;
	bool all_zero = true;
#pragma unroll
	for (int c = 0; c < 3; c++) all_zero = all_zero && box[c * Box::VOXELS + at] == 0.0f;
	acc[0] = acc[1] = acc[2] = 0.0f;
	if (all_zero) return;
#pragma unroll
	for (int c = 0; c < 3; c++) {
#pragma unroll
		for (int t = 0; t < 2 * R + 1; t++) acc[c] += box[c * Box::VOXELS + at + (t - R) * Box::STRIDE] * k[t];
	}
}

template<int R, int AXIS>
__device__ __forceinline__ int slav_staged_filter_offset(int lx, int ly, int lz) {
	typedef SlavPassBox<R, AXIS> Box;
	return ((lx + Box::LO_X) * Box::Y + (ly + Box::LO_Y)) * Box::Z + (lz + Box::LO_Z);
}

constexpr int SLAV_FILTER_STAGES = 4, SLAV_RESAMPLE_STAGES = 3;

template<int R, int AXIS>
static __global__ void __launch_bounds__(SLAV_CONSUMERS + 32, 1) k_slav_brick_filter_tma(SlavFilterArgs a, SlavBrickArgs b,
		const __grid_constant__ CUtensorMap in_map) {
	typedef SlavPassBox<R, AXIS> Box;
	constexpr int BOX_BYTES = 3 * Box::VOXELS * 4, STAGE_BYTES = BOX_BYTES + SLAV_LIST_BYTES;
	if (a.status[a.iteration]) return;
	extern __shared__ __align__(128) unsigned char slav_smem[];
	__shared__ SlavPipe<SLAV_FILTER_STAGES> pipe;
	slav_pipe_init(pipe, SLAV_CONSUMERS / 32);
	if (threadIdx.x >= SLAV_CONSUMERS) {
		if (threadIdx.x == SLAV_CONSUMERS) {
			const int total = *b.active_count;
			for (int n = 0;; n++) {
				const int s = n % SLAV_FILTER_STAGES;
				unsigned char* stage = slav_smem + s * STAGE_BYTES;
				const int brick = slav_pipe_claim(pipe, b, total, n, reinterpret_cast<unsigned short*>(stage + BOX_BYTES), BOX_BYTES);
				if (brick < 0) break;
				int x0, y0, z0;
				slav_brick_origin(b, brick, x0, y0, z0);
				tma_load_4d(stage, &in_map, z0 - Box::LO_Z, y0 - Box::LO_Y, x0 - Box::LO_X, 0, &pipe.full[s]);
			}
		}
		return;
	}
	const int N = (int) a.g.N;
	for (int n = 0;; n++) {
		const int s = n % SLAV_FILTER_STAGES;
		mbar_wait(&pipe.full[s], (n / SLAV_FILTER_STAGES) & 1);
		const int brick = pipe.brick[s];
		if (brick < 0) break;
		const unsigned char* stage = slav_smem + s * STAGE_BYTES;
		const float* box = reinterpret_cast<const float*>(stage);
		const unsigned short* list = reinterpret_cast<const unsigned short*>(stage + BOX_BYTES);
		int x0, y0, z0;
		slav_brick_origin(b, brick, x0, y0, z0);
		const int count = pipe.count[s];
		for (int j = threadIdx.x; j < count; j += SLAV_CONSUMERS) {
			const int local = list[j];
			const int lx = local >> 8, ly = (local >> 5) & 7, lz = local & 31;
			const int idx = ((x0 + lx) * a.g.n[1] + y0 + ly) * a.g.n[2] + z0 + lz;
			float acc[3];
			slav_staged_filter_voxel<R, AXIS>(box, slav_staged_filter_offset<R, AXIS>(lx, ly, lz), a.k, acc);
#pragma unroll
			for (int c = 0; c < 3; c++) a.out[c * N + idx] = acc[c];
		}
		slav_pipe_release(pipe, s);
	}
}

// axis-2 pass from the staged box + masked re-warp (taps from the staged live box) + maximum warp length
// stage: pass input [3][8][8][40], live [10][10][40], canonical [8][8][32], list
template<int R>
static __global__ void __launch_bounds__(SLAV_CONSUMERS + 32, 1) k_slav_brick_filter_resample_tma(SlavFilterArgs a,
		SlavResampleArgs ra, SlavBrickArgs b, const __grid_constant__ CUtensorMap in_map,
		const __grid_constant__ CUtensorMap live_map, const __grid_constant__ CUtensorMap canonical_map) {
	typedef SlavPassBox<R, 2> Box;
	constexpr int BOX_BYTES = (3 * Box::VOXELS + SLAV_TERMS_TILE + SLAV_BRICK_VOXELS) * 4, STAGE_BYTES = BOX_BYTES + SLAV_LIST_BYTES;
	if (a.status[a.iteration]) return;
	extern __shared__ __align__(128) unsigned char slav_smem[];
	__shared__ SlavPipe<SLAV_RESAMPLE_STAGES> pipe;
	slav_pipe_init(pipe, SLAV_CONSUMERS / 32);
	float sq_report = 0.0f;
	if (threadIdx.x >= SLAV_CONSUMERS) {
		if (threadIdx.x == SLAV_CONSUMERS) {
			const int total = *b.active_count;
			for (int n = 0;; n++) {
				const int s = n % SLAV_RESAMPLE_STAGES;
				unsigned char* stage = slav_smem + s * STAGE_BYTES;
				const int brick = slav_pipe_claim(pipe, b, total, n, reinterpret_cast<unsigned short*>(stage + BOX_BYTES), BOX_BYTES);
				if (brick < 0) break;
				int x0, y0, z0;
				slav_brick_origin(b, brick, x0, y0, z0);
				float* box = reinterpret_cast<float*>(stage);
				tma_load_4d(box, &in_map, z0 - Box::LO_Z, y0, x0, 0, &pipe.full[s]);
				tma_load_4d(box + 3 * Box::VOXELS, &live_map, z0 - 4, y0 - 1, x0 - 1, 0, &pipe.full[s]);
				tma_load_4d(box + 3 * Box::VOXELS + SLAV_TERMS_TILE, &canonical_map, z0, y0, x0, 0, &pipe.full[s]);
			}
		}
	} else {
		const int N = (int) a.g.N;
		for (int n = 0;; n++) {
			const int s = n % SLAV_RESAMPLE_STAGES;
			mbar_wait(&pipe.full[s], (n / SLAV_RESAMPLE_STAGES) & 1);
			const int brick = pipe.brick[s];
			if (brick < 0) break;
			const unsigned char* stage = slav_smem + s * STAGE_BYTES;
			const float* box = reinterpret_cast<const float*>(stage);
			const float* live_tile = box + 3 * Box::VOXELS;
			const float* canonical_box = live_tile + SLAV_TERMS_TILE;
			const unsigned short* list = reinterpret_cast<const unsigned short*>(stage + BOX_BYTES);
			int x0, y0, z0;
			slav_brick_origin(b, brick, x0, y0, z0);
			SlavLiveTile tile;
			tile.data = live_tile;
			tile.lo[0] = x0 - 1;
			tile.lo[1] = y0 - 1;
			tile.lo[2] = z0 - 4;
			tile.ext[0] = SLAV_BRICK_X + 2;
			tile.ext[1] = SLAV_BRICK_Y + 2;
			tile.ext[2] = SLAV_STAGE_Z;
			const int count = pipe.count[s];
			for (int j = threadIdx.x; j < count; j += SLAV_CONSUMERS) {
				const int local = list[j];
				const int lx = local >> 8, ly = (local >> 5) & 7, lz = local & 31;
				const int q[3] = { x0 + lx, y0 + ly, z0 + lz };
				const int idx = (q[0] * a.g.n[1] + q[1]) * a.g.n[2] + q[2];
				const float canonical_value = canonical_box[local];
				float update[3], w[3], new_value;
				slav_staged_filter_voxel<R, 2>(box, slav_staged_filter_offset<R, 2>(lx, ly, lz), a.k, update);
				const float live_value = live_tile[((lx + 1) * (SLAV_BRICK_Y + 2) + (ly + 1)) * SLAV_STAGE_Z + (lz + 4)];
				slav_resample_voxel<3, true>(ra, idx, update, live_value, canonical_value, new_value, w, sq_report, &tile, q);
				ra.new_live[idx] = new_value;
#pragma unroll
				for (int c = 0; c < 3; c++) ra.warp[c * N + idx] = w[c];
				if (slav_truncated(new_value) && slav_truncated(canonical_value)) b.leave_list[atomicAdd(b.leave_count, 1)] = idx;
			}
			slav_pipe_release(pipe, s);
		}
	}
	if (ra.max_sq_bits != nullptr) block_atomic_max(sq_report, ra.max_sq_bits);
}

inline unsigned slav_pipe_blocks() {
	static int sms = 0;
	if (sms == 0) {
		int device = 0;
		cudaGetDevice(&device);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
		if (sms <= 0) sms = 148;
	}
	return (unsigned) sms;
}

inline void launch_slav_brick_terms_tma(const SlavGradientArgs& ga, const SlavBrickArgs& brick, const SlavBrickMaps& maps,
		int live_parity, cudaStream_t stream) {
	constexpr int bytes = SLAV_TERMS_STAGES * SLAV_TERMS_STAGE_BYTES;
	static const bool configured = cudaFuncSetAttribute(k_slav_brick_terms_tma, cudaFuncAttributeMaxDynamicSharedMemorySize,
			bytes) == cudaSuccess;
	(void) configured;
	k_slav_brick_terms_tma<<<counted(slav_pipe_blocks()), SLAV_CONSUMERS + 32, bytes, stream>>>(ga, brick, maps.live[live_parity],
			maps.warp, maps.canonical);
}

template<int R, int AXIS>
inline void launch_slav_brick_filter_tma(const SlavFilterArgs& fa, const SlavBrickArgs& brick, const CUtensorMap& map,
		cudaStream_t stream) {
	constexpr int bytes = SLAV_FILTER_STAGES * (3 * SlavPassBox<R, AXIS>::VOXELS * 4 + SLAV_LIST_BYTES);
	static const bool configured = cudaFuncSetAttribute(k_slav_brick_filter_tma<R, AXIS>,
			cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
	(void) configured;
	k_slav_brick_filter_tma<R, AXIS> <<<counted(slav_pipe_blocks()), SLAV_CONSUMERS + 32, bytes, stream>>>(fa, brick, map);
}

template<int R>
inline void launch_slav_brick_filter_resample_tma(const SlavFilterArgs& fa, const SlavResampleArgs& ra,
		const SlavBrickArgs& brick, const SlavBrickMaps& maps, int live_parity, cudaStream_t stream) {
	constexpr int bytes = SLAV_RESAMPLE_STAGES
			* ((3 * SlavPassBox<R, 2>::VOXELS + SLAV_TERMS_TILE + SLAV_BRICK_VOXELS) * 4 + SLAV_LIST_BYTES);
	static const bool configured = cudaFuncSetAttribute(k_slav_brick_filter_resample_tma<R>,
			cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
	(void) configured;
	k_slav_brick_filter_resample_tma<R> <<<counted(slav_pipe_blocks()), SLAV_CONSUMERS + 32, bytes, stream>>>(fa, ra, brick,
			maps.pass[2], maps.live[live_parity], maps.canonical);
}

#endif  // __CUDACC__

}  // namespace lsf
