// kernels3d_tma.cuh -- third generation of the 3D hierarchical iteration (sm_100a): TMA-fed stage 1 and a
// y-marching filter kernel.
//
// ncu of the second generation (capture not kept; figures quoted in profiles/r1_ncu_tma_v3.md) showed both kernels to be limited by instruction issue,
// not by HBM: 452 + 255 executed thread-instructions per voxel, a quarter of them 64-bit address arithmetic for
// global loads (IADD3/IADD3.X/LEA/LEA.HI.X per distinct address), the rest inflated by branch-free border selects and
// by one block-wide barrier per filter pass. This generation removes that overhead instead of re-tiling it:
//
//   k_hier_stage1_tma<TIKHONOV,R,NS>  stage 1 (gather + data term + Tikhonov term, reference optimizer.tpp:186-200) and
//                                     the axis-0 filter pass (convolution.cpp:240-267). The warp planes, the canonical
//                                     field and the halo'd tile of the previous gradient arrive in shared memory by TMA
//                                     (cp.async.bulk.tensor, one elected thread, mbarrier completion, NS-deep ring along
//                                     the marching axis); out-of-volume halo elements are zero-filled by the TMA unit.
//                                     Every stencil operand is then an LDS with an immediate offset from one base
//                                     register. The only global loads left are the eight 128-bit taps of the trilinear
//                                     gather. Interior voxels take a border-free Laplacian path (uniform branch).
//                                     The trilinear blend uses packed f32x2 multiplies/adds (IEEE round-to-nearest per
//                                     lane, no contraction: bit-identical to the scalar sequence).
//   k_sobolev_ymarch<R>               axis-1 and axis-2 passes + warp update + max-norm (convolution.cpp:268-331,
//                                     optimizer.tpp:207-211). A thread owns one z column of one x plane and marches
//                                     along y: the axis-1 pass is a register chain of partial sums (no halo, no shared
//                                     memory), the axis-2 pass reads a double-buffered shared row (one barrier per row).
//
// HBM traffic is unchanged (56 + 48 B/voxel); arithmetic is unchanged (float32, reference order, no FMA).
#pragma once

#include "kernels3d_split.cuh"

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

namespace lsf {

// ---------------------------------------------------------------------------------------------- host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
		const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
		CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
	static EncodeTiledFn fn = nullptr;
	if (fn == nullptr) {
		void* p = nullptr;
		cudaDriverEntryPointQueryResult status;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &status) == cudaSuccess
				&& status == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
	}
	return fn;
}

// Tensor map over `channels` planes of an [X][Y][Z] float field (plane c starts at base + c * channel_stride);
// box = box_z x box_y voxels of one x plane, all channels. Out-of-bounds elements are filled with zeros.
inline int make_planes_map(CUtensorMap* map, const float* base, int channels, long long channel_stride, const Grid3& g,
		int box_z, int box_y) {
	EncodeTiledFn encode = encode_tiled_fn();
	LSF_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
	const cuuint64_t dims[4] = { (cuuint64_t) g.Z, (cuuint64_t) g.Y, (cuuint64_t) g.X, (cuuint64_t) channels };
	const cuuint64_t strides[3] = { (cuuint64_t) g.Z * 4, (cuuint64_t) g.Y * g.Z * 4, (cuuint64_t) channel_stride * 4 };
	const cuuint32_t box[4] = { (cuuint32_t) box_z, (cuuint32_t) box_y, 1u, (cuuint32_t) channels };
	const cuuint32_t element_strides[4] = { 1, 1, 1, 1 };
	const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box,
			element_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
			CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	LSF_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d (dims %d x %d x %d, box %d x %d)", (int) r,
			g.X, g.Y, g.Z, box_z, box_y);
	return LSF_OK;
}

// Tensor map over the padded pack of a level ([X + 4][Y + 4][Z + 4] float4 entries, seen as float32 rows of 4 (Z + 4)
// elements); box = one plane of box_entries x box_rows entries. Only used for L2 prefetches.
inline int make_pack_map(CUtensorMap* map, const float4* pack, int planes, const Grid3& g, int box_entries, int box_rows) {
	EncodeTiledFn encode = encode_tiled_fn();
	LSF_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
	const cuuint64_t dims[3] = { (cuuint64_t) (g.Z + 4) * 4, (cuuint64_t) g.Y + 4, (cuuint64_t) planes + 4 };
	const cuuint64_t strides[2] = { (cuuint64_t) (g.Z + 4) * 16, (cuuint64_t) (g.Y + 4) * (g.Z + 4) * 16 };
	const cuuint32_t box[3] = { (cuuint32_t) box_entries * 4, (cuuint32_t) box_rows, 1u };
	const cuuint32_t element_strides[3] = { 1, 1, 1 };
	const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float4*>(pack), dims, strides, box,
			element_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
			CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	LSF_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d for the pack (dims %d x %d x %d)", (int) r,
			planes, g.Y, g.Z);
	return LSF_OK;
}

#ifdef __CUDACC__

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
	return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	const uint32_t addr = smem_addr(bar);
	uint32_t done;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
				: "=r"(done) : "r"(addr), "r"(parity) : "memory");
	} while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// one box of a 4-D tensor map -> shared memory, completion on `bar`
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
		uint64_t* bar) {
	asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
			::"r"(smem_addr(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
			: "memory");
}
// hint: bring one box of a 3-D tensor map into L2 (no shared-memory destination, no completion)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
	asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
			::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
	float v;
	asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
	return v;
}

// packed float32 pairs: mul.rn.f32x2 / add.rn.f32x2 round each lane like the scalar instructions
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
	f32x2 r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
	asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
	f32x2 r;
	asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}
// a + b as fma(a, one, b) with one = {1.0f, 1.0f} supplied at run time (a kernel argument): a * 1 + b is exactly the
// rounded sum. ptxas 12.9 contracts mul.rn.f32x2 followed by add.rn.f32x2 into FFMA2 even under --fmad=false (and
// sees through a literal 1.0), which would change the reference's rounding; an opaque multiplier cannot be folded.
constexpr unsigned long long F32X2_ONE = 0x3f8000003f800000ull;
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b, f32x2 one) {
	f32x2 r;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(one), "l"(b));
	return r;
}
// v0 * i + v1 * r on all four lanes of a pack entry (two f32x2 halves)
__device__ __forceinline__ ulonglong2 blend4(ulonglong2 v0, ulonglong2 v1, f32x2 i, f32x2 r, f32x2 one) {
	ulonglong2 o;
	o.x = add2(mul2(v0.x, i), mul2(v1.x, r), one);
	o.y = add2(mul2(v0.y, i), mul2(v1.y, r), one);
	return o;
}

// trilinear gather (same arithmetic and citations as gather4i) with packed blends
__device__ __forceinline__ float4 gather4p(const float4* __restrict__ pack, int X, int Y, int Z, int x, int y, int z,
		float wx, float wy, float wz, f32x2 one) {
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
	const ulonglong2* p = reinterpret_cast<const ulonglong2*>(pack) + ((bx + 2) * sx + (by + 2) * sy + (bz + 2));
	// plane bx + 1 first: it is the one this x step touches for the first time (the likely L1 misses)
	const ulonglong2 v100 = __ldg(p + sx), v101 = __ldg(p + sx + 1);
	const ulonglong2 v110 = __ldg(p + sx + sy), v111 = __ldg(p + sx + sy + 1);
	const ulonglong2 v000 = __ldg(p), v001 = __ldg(p + 1);
	const ulonglong2 v010 = __ldg(p + sy), v011 = __ldg(p + sy + 1);
	const f32x2 izz = pack2(iz, iz), rzz = pack2(rz, rz);
	const ulonglong2 i00 = blend4(v000, v001, izz, rzz, one);
	const ulonglong2 i01 = blend4(v010, v011, izz, rzz, one);
	const ulonglong2 i10 = blend4(v100, v101, izz, rzz, one);
	const ulonglong2 i11 = blend4(v110, v111, izz, rzz, one);
	const f32x2 iyy = pack2(iy, iy), ryy = pack2(ry, ry);
	const ulonglong2 i0 = blend4(i00, i01, iyy, ryy, one);
	const ulonglong2 i1 = blend4(i10, i11, iyy, ryy, one);
	const ulonglong2 o = blend4(i0, i1, pack2(ix, ix), pack2(rx, rx), one);
	float4 s;
	unpack2(o.x, s.x, s.y);
	unpack2(o.y, s.z, s.w);
	return s;
}

// trilinear gather with packed blends (arithmetic and citations: gather4i / gather4i_slab); the voxel's coordinates
// arrive as floats (hoisted conversions)
template<bool SLAB>
__device__ __forceinline__ float4 gather4q(const HierIterArgs& a, int sx, int sy, float fx, float fy, float fz, float wx,
		float wy, float wz, f32x2 one) {
	const float lookup_x = fx + wx;
	const float lookup_y = fy + wy;
	const float lookup_z = fz + wz;
	int bx = __float2int_rd(lookup_x);
	int by = __float2int_rd(lookup_y);
	int bz = __float2int_rd(lookup_z);
	const float rx = lookup_x - (float) bx, ry = lookup_y - (float) by, rz = lookup_z - (float) bz;
	const float ix = 1.0f - rx, iy = 1.0f - ry, iz = 1.0f - rz;
	if (SLAB) {
		bx = min(max(bx, -2), a.X_global) - a.pack_origin;
		if ((a.pack_interior_low && bx < 0) || (a.pack_interior_high && bx + 1 > a.pack_X - 1)) {
			if (a.violation != nullptr) *a.violation = 1;
		}
		bx = min(max(bx, -2), a.pack_X);
	} else {
		bx = min(max(bx, -2), a.g.X);
	}
	by = min(max(by, -2), a.g.Y);
	bz = min(max(bz, -2), a.g.Z);
	const ulonglong2* p = reinterpret_cast<const ulonglong2*>(a.pack) + ((bx + 2) * sx + (by + 2) * sy + (bz + 2));
	const ulonglong2 v000 = __ldg(p), v001 = __ldg(p + 1);
	const ulonglong2 v010 = __ldg(p + sy), v011 = __ldg(p + sy + 1);
	const ulonglong2 v100 = __ldg(p + sx), v101 = __ldg(p + sx + 1);
	const ulonglong2 v110 = __ldg(p + sx + sy), v111 = __ldg(p + sx + sy + 1);
	const f32x2 izz = pack2(iz, iz), rzz = pack2(rz, rz);
	const ulonglong2 i00 = blend4(v000, v001, izz, rzz, one);
	const ulonglong2 i01 = blend4(v010, v011, izz, rzz, one);
	const ulonglong2 i10 = blend4(v100, v101, izz, rzz, one);
	const ulonglong2 i11 = blend4(v110, v111, izz, rzz, one);
	const f32x2 iyy = pack2(iy, iy), ryy = pack2(ry, ry);
	const ulonglong2 i0 = blend4(i00, i01, iyy, ryy, one);
	const ulonglong2 i1 = blend4(i10, i11, iyy, ryy, one);
	const ulonglong2 o = blend4(i0, i1, pack2(ix, ix), pack2(rx, rx), one);
	float4 s;
	unpack2(o.x, s.x, s.y);
	unpack2(o.y, s.z, s.w);
	return s;
}

// ---------------------------------------------------------------------------------------------- stage 1 + axis-0 pass
template<bool TIKHONOV>
struct Stage1Tile {
	static constexpr int TZ = 32, TY = 8;             // voxels per block and plane (threads)
	static constexpr int GZ = TZ + 8, GY = TY + 2;    // g_prev box: 1-voxel halo; along z the box starts 4 voxels early
	static constexpr int GZ0 = 4;                     // because a TMA box must start on a 16-byte boundary (measured:
	                                                  // a start coordinate of z0 - 1 raises "illegal instruction")
	static constexpr int GP_TX = TIKHONOV ? 3 * GY * GZ * 4 : 0;  // bytes the TMA unit reports
	static constexpr int GP_BYTES = (GP_TX + 127) / 128 * 128;
	static constexpr int WP_BYTES = 3 * TY * TZ * 4;
	static constexpr int CN_BYTES = TY * TZ * 4;
	static constexpr int STAGE_BYTES = GP_BYTES + WP_BYTES + CN_BYTES;
	static constexpr int STAGE_TX = GP_TX + WP_BYTES + CN_BYTES;
	// L2 prefetch box of the pack: the tile's gather footprint (TZ + 1 entries x TY + 1 rows) with a margin either side
	static constexpr int PACK_BOX_Z0 = 6, PACK_BOX_Y0 = 3;
	static constexpr int PACK_BOX_Z = TZ + 1 + 2 * PACK_BOX_Z0 + 3, PACK_BOX_Y = TY + 1 + 2 * PACK_BOX_Y0;  // 48 x 15
};

// APPLY (deferred warp update, needs TIKHONOV): the previous iteration's filter kernel wrote only its gradient; this
// kernel reads the warp BEFORE that update from a.warp, applies `warp - g_prev * rate` (reference optimizer.tpp:207,
// the same two roundings, one kernel later) for the gather and writes the updated warp of its own planes to
// a.warp_out (a different buffer: neighbouring x-chunks still read the old planes). One warp read less per iteration.
// SYM: the filter kernel is symmetric (k[q] == k[K-1-q] bit for bit, true of every Sobolev kernel): the axis-0 chain
// multiplies each gradient once per distinct tap (the two products are the same rounded value).
// EXACT: one whole volume (no batch, no slab) whose planes the tiles cover exactly (Y % TY == 0, Z % TZ == 0): no thread
// is outside the volume, the x-chunks come from XPassArgs::chunk_bounds.
template<bool TIKHONOV, int R, int NS, bool DEC, bool FUSE = false, int PD = 0, bool APPLY = false, bool SLAB = false,
		bool SYM = false, bool EXACT = false>
static __global__ void __launch_bounds__(256, 3) k_hier_stage1_tma(const __grid_constant__ CUtensorMap map_g,
		const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_c,
		const __grid_constant__ CUtensorMap map_p, HierIterArgs a, XPassArgs t) {
	typedef Stage1Tile<TIKHONOV> T;
	constexpr int K = 2 * R + 1;
	// batch of pairs (HierIterArgs::batch_X): this block's pair owns planes [x_lo, x_hi) of the allocation
	int x_lo = 0, x_hi = a.g.X, chunk_index = blockIdx.z;
	const float4* pack = a.pack;
	unsigned* slots = a.max_sq_bits;
	if (!SLAB && !EXACT && a.batch_X > 0) {
		const int pair = blockIdx.z / a.batch_chunks;
		chunk_index = blockIdx.z - pair * a.batch_chunks;
		x_lo = pair * a.batch_X;
		x_hi = x_lo + a.batch_X;
		pack += pair * a.batch_pack_stride;
		if (slots != nullptr) slots += pair * a.batch_slot_stride;
	}
	pdl_launch_dependents();
	extern __shared__ __align__(128) unsigned char stage_memory[];
	__shared__ uint64_t full[NS];
	__shared__ uint64_t empty[NS];  // DEC: one arrival per warp once it has read the slot

	const int X = x_hi, Y = a.g.Y, Z = a.g.Z;  // X: end of this volume's planes in the allocation
	const int tz = threadIdx.x, ty = threadIdx.y;
	const int z0 = blockIdx.x * T::TZ, y0 = blockIdx.y * T::TY;
	const int z = z0 + tz, y = y0 + ty;
	const bool valid = EXACT || (z < Z && y < Y);
	const int YZ = Y * Z;
	const int N = (int) a.g.N;
	// SLAB (slab decomposition, slab.py): the fields are an allocation of X planes of which [x_begin, x_end) are processed;
	// allocation plane p is plane p + x_origin of a level of X_global planes (border rules, gather look-ups)
	// EXACT kernels (one whole volume) take their chunk from the table: chunks may have unequal lengths
	const int xs = EXACT ? t.chunk_bounds[chunk_index] : (SLAB ? a.x_begin : x_lo) + chunk_index * t.x_chunk;
	const int xe = EXACT ? t.chunk_bounds[chunk_index + 1] : min(SLAB ? a.x_end : X, xs + t.x_chunk);
	const int origin = SLAB ? a.x_origin : -x_lo, Xg = SLAB ? a.X_global : X - x_lo;
	// FUSE (no Sobolev kernel configured): no filter pass, the warp update and the max-norm (reference
	// optimizer.tpp:207-211) happen here and the kernel is the whole iteration
	const int x_first = FUSE ? xs : max(xs - R, x_lo);  // planes below the volume contribute zeros to accumulators that are still zero
	const int x_stop = FUSE ? xe : xe + R;           // planes >= X contribute zeros
	const int p_last = min(x_stop, X - 1);  // last plane fetched (the Tikhonov term looks one plane ahead)
	const bool leader = tz == 0 && ty == 0;
	float best = 0.0f;

	if (leader) {
#pragma unroll
		for (int s = 0; s < NS; s++) {
			mbar_init(&full[s], 1);
			mbar_init(&empty[s], T::TY);
		}
		mbar_fence_init();
	}
	__syncthreads();
	pdl_wait();  // everything above is independent of the previous kernel's output
	if (a.check_convergence && level_converged(slots, a.iteration, a.threshold)) return;
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

	// this thread's byte offsets inside a stage
	const uint32_t stage_base = smem_addr(stage_memory);
	const uint32_t off_g = ((ty + 1) * T::GZ + tz + T::GZ0) * 4;           // centre of component 0 in the halo'd tile
	const uint32_t off_w = T::GP_BYTES + (ty * T::TZ + tz) * 4;
	const uint32_t off_c = T::GP_BYTES + T::WP_BYTES + (ty * T::TZ + tz) * 4;
	constexpr uint32_t G_COMP = T::GY * T::GZ * 4, W_COMP = T::TY * T::TZ * 4, G_ROW = T::GZ * 4;

	// axis-0 chain: components 0 and 1 as one packed f32x2 accumulator per tap, component 2 scalar
	f32x2 acc01[K];
	float acc2[K];
#pragma unroll
	for (int q = 0; q < K; q++) {
		acc01[q] = 0ull;
		acc2[q] = 0.0f;
	}

	// the Laplacian's border rules (reference gradients.tpp:28-35) apply on the faces of the volume only
	const bool yz_border = y0 == 0 || y0 + T::TY >= Y || z0 == 0 || z0 + T::TZ >= Z;
	const int kind_y = border_kind(y, Y), kind_z = border_kind(z, Z);

	int slot = 0;
	uint32_t phase = 0;
	mbar_wait(&full[0], 0);
	float prev[3] = { 0.f, 0.f, 0.f }, cur[3] = { 0.f, 0.f, 0.f };
	if (TIKHONOV) {
#pragma unroll
		for (int c = 0; c < 3; c++) {
			cur[c] = lds_f32(stage_base + off_g + c * G_COMP);
			if (x_first > x_lo && valid) prev[c] = __ldg(a.g_prev + c * N + (x_first - 1) * YZ + y * Z + z);
		}
	}
	int out = (x_first - R) * YZ + y * Z + z;  // index of the plane completed by the current step
#pragma unroll 1
	for (int x = x_first; x < x_stop; x++, out += YZ) {
		float g[3] = { 0.f, 0.f, 0.f };
		if (x < X) {
			const uint32_t st = stage_base + slot * T::STAGE_BYTES;
			const int next_slot = slot + 1 == NS ? 0 : slot + 1;
			mbar_wait(&full[slot], phase);
			float wx = lds_f32(st + off_w), wy = lds_f32(st + off_w + W_COMP), wz = lds_f32(st + off_w + 2 * W_COMP);
			const float cn = lds_f32(st + off_c);
			if (APPLY) {
				wx = wx - cur[0] * a.rate;  // cur[] is the previous gradient at this voxel
				wy = wy - cur[1] * a.rate;
				wz = wz - cur[2] * a.rate;
				if (valid && x >= xs && x < xe) {
					const int at = out + R * YZ;
					a.warp_out[at] = wx;
					a.warp_out[N + at] = wy;
					a.warp_out[2 * N + at] = wz;
				}
			}
			float lap[3] = { 0.f, 0.f, 0.f };
			if (TIKHONOV) {
				const bool has_next = x + origin + 1 < Xg;
				if (has_next) mbar_wait(&full[next_slot], next_slot == 0 ? phase ^ 1u : phase);
				const uint32_t nst = stage_base + next_slot * T::STAGE_BYTES;
				if (yz_border || x + origin == 0 || !has_next) {
					const int kind_x = border_kind(x + origin, Xg);
#pragma unroll
					for (int c = 0; c < 3; c++) {
						const uint32_t p = st + off_g + c * G_COMP;
						const float next = has_next ? lds_f32(nst + off_g + c * G_COMP) : 0.0f;
						const float ym = lds_f32(p - G_ROW), yp = lds_f32(p + G_ROW);
						const float zm = lds_f32(p - 4), zp = lds_f32(p + 4);
						float l = laplace_select(prev[c], cur[c], next, kind_x);
						l += laplace_select(ym, cur[c], yp, kind_y);
						l += laplace_select(zm, cur[c], zp, kind_z);
						lap[c] = l;
						prev[c] = cur[c];
						cur[c] = next;
					}
				} else {
#pragma unroll
					for (int c = 0; c < 3; c++) {
						const uint32_t p = st + off_g + c * G_COMP;
						const float next = lds_f32(nst + off_g + c * G_COMP);
						const float ym = lds_f32(p - G_ROW), yp = lds_f32(p + G_ROW);
						const float zm = lds_f32(p - 4), zp = lds_f32(p + 4);
						const float twice = 2.0f * cur[c];
						float l = (next - twice) + prev[c];
						l += (yp - twice) + ym;
						l += (zp - twice) + zm;
						lap[c] = l;
						prev[c] = cur[c];
						cur[c] = next;
					}
				}
			}
			if (DEC) {
				// this warp (one row of the tile) has read everything it needs from `slot`; the leader refills the slot of
				// the previous plane as soon as all warps have released it -- no block-wide barrier, the warps of a block
				// drift up to NS - 1 planes apart and their gather latencies overlap
				__syncwarp();
				if (tz == 0) mbar_arrive(&empty[slot]);
				if (leader && x > x_first) {
					const int refill = slot == 0 ? NS - 1 : slot - 1;
					mbar_wait(&empty[refill], slot == 0 ? phase ^ 1u : phase);
					if (x - 1 + NS <= p_last) fetch(x - 1 + NS, refill);
				}
				if (PD > 0 && leader && x + PD < min(x_stop, X)) {
					// L2 prefetch of the pack plane that the gather of plane x + PD touches for the first time, placed by the
					// displacement of the tile's first voxel (a hint: the warp field is smooth, the box has a margin)
					int s2 = slot + PD;
					uint32_t phase2 = phase;
					if (s2 >= NS) {
						s2 -= NS;
						phase2 ^= 1u;
					}
					mbar_wait(&full[s2], phase2);
					const uint32_t st2 = stage_base + s2 * T::STAGE_BYTES + T::GP_BYTES;
					const int bx = min(max(__float2int_rd((float) (x + PD) + lds_f32(st2)), -2), X) + 1;
					const int by = min(max(__float2int_rd((float) y0 + lds_f32(st2 + W_COMP)), -2), Y);
					const int bz = min(max(__float2int_rd((float) z0 + lds_f32(st2 + 2 * W_COMP)), -2), Z);
					tma_prefetch_3d(&map_p, 4 * (bz + 2 - T::PACK_BOX_Z0), by + 2 - T::PACK_BOX_Y0, bx + 2);
				}
			}
			const float4 s = SLAB ? gather4q<true>(a, (Y + 4) * (Z + 4), Z + 4, (float) (x + origin), (float) y, (float) z, wx, wy,
					wz, t.one2) : gather4p(pack, Xg, Y, Z, x + origin, y, z, wx, wy, wz, t.one2);
			const float diff = s.x - cn;
			g[0] = (s.y * diff) * a.amplifier;
			g[1] = (s.z * diff) * a.amplifier;
			g[2] = (s.w * diff) * a.amplifier;
			if (TIKHONOV) {
				g[0] = g[0] - lap[0] * a.strength;
				g[1] = g[1] - lap[1] * a.strength;
				g[2] = g[2] - lap[2] * a.strength;
			}
			if (FUSE && valid) {
				const int at = out + R * YZ;  // this plane
				if (a.g_out != nullptr) {
					a.g_out[at] = g[0];
					a.g_out[N + at] = g[1];
					a.g_out[2 * N + at] = g[2];
				}
				if (a.warp_out != nullptr) {  // nullptr: gradient only (slab mode, the filter phase updates the warp)
					a.warp_out[at] = wx - g[0] * a.rate;
					a.warp_out[N + at] = wy - g[1] * a.rate;
					a.warp_out[2 * N + at] = wz - g[2] * a.rate;
					float sq = g[0] * g[0];
					sq += g[1] * g[1];
					sq += g[2] * g[2];
					if (sq > best) best = sq;
				}
			}
		}
		if (!FUSE) {
			// axis-0 filter pass: plane x is tap q of output plane x + R - q; acc[c][q] holds the partial sum (taps 0..q)
			// of output x + R - q, so adding in place from the oldest output down reproduces sum_{q ascending} in[.]*k[q]
			const f32x2 g01 = pack2(g[0], g[1]);
			f32x2 p01[K];
			float p2[K];
#pragma unroll
			for (int q = 0; q < K; q++) {
				p01[q] = (SYM && q > R) ? p01[K - 1 - q] : mul2(g01, t.k2[q]);
				p2[q] = (SYM && q > R) ? p2[K - 1 - q] : g[2] * t.k[q];
			}
#pragma unroll
			for (int q = K - 1; q >= 1; q--) {
				acc01[q] = add2(acc01[q - 1], p01[q], t.one2);
				acc2[q] = acc2[q - 1] + p2[q];
			}
			acc01[0] = p01[0];
			acc2[0] = p2[0];
			if (x - R >= xs && valid) {
				float h0, h1;
				unpack2(acc01[K - 1], h0, h1);
				a.g_out[out] = h0;
				a.g_out[N + out] = h1;
				a.g_out[2 * N + out] = acc2[K - 1];
			}
		}
		if (!DEC) {
			// every thread has read stage `slot`: refill it with the plane NS steps ahead
			__syncthreads();
			if (leader && x + NS <= p_last) fetch(x + NS, slot);
		}
		slot++;
		if (slot == NS) {
			slot = 0;
			phase ^= 1u;
		}
	}
	if (FUSE && a.warp_out != nullptr) block_atomic_max(best, slots + a.iteration);
}

// ---------------------------------------------------------------------------------------------- axis-1 / axis-2 passes
struct YMarchArgs {
	const float* in;   // planes after the axis-0 pass
	float* out;        // planes: filtered gradient (nullptr when nothing reads it)
	float* warp;       // planes, updated in place: warp -= out * rate
	Grid3 g;
	float k[7];        // flipped taps: k[q] multiplies in[i - R + q]
	float rate, threshold;
	unsigned* max_sq_bits;
	int iteration;
	int check_convergence;
	int y_chunk;       // output rows per block
	int x_begin;       // first plane (blockIdx.y counts from here)
	int tile_z;        // output columns per block; blockDim.x = tile_z (+ 32 halo-column threads when gridDim.x > 1)
};

template<int R>
static __global__ void __launch_bounds__(288) k_sobolev_ymarch(YMarchArgs a) {
	constexpr int K = 2 * R + 1;
	if (a.check_convergence && level_converged(a.max_sq_bits, a.iteration, a.threshold)) return;
	extern __shared__ float row_memory[];  // [2][3][tile_z + 2R]
	const int Y = a.g.Y, Z = a.g.Z;
	const int NT = a.tile_z;
	const int W = NT + 2 * R;
	const int tid = threadIdx.x;
	const int z0 = blockIdx.x * NT;
	const int x = a.x_begin + blockIdx.y;
	const int ys = blockIdx.z * a.y_chunk;
	const int ye = min(Y, ys + a.y_chunk);
	// row-buffer slot of this thread: owners hold R .. R+NT-1, the halo threads the R columns either side
	const bool owner = tid < NT;
	const int j = tid - NT;
	const int zl = owner ? tid + R : (j < R ? j : NT + j);
	const int z = z0 - R + zl;
	const bool active = (owner || j < 2 * R) && z >= 0 && z < Z;
	const bool writes = owner && active;
	for (int i = tid; i < 6 * W; i += blockDim.x) row_memory[i] = 0.0f;  // columns outside the volume stay zero
	__syncthreads();

	float acc[3][K];
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int q = 0; q < K; q++) acc[c][q] = 0.0f;
	const int r_first = max(ys - R, 0);
	const int r_stop = ye + R;
	const int r_load_end = min(r_stop, Y);
	const int N = (int) a.g.N;
	int at = (x * Y + r_first) * Z + z;  // row being consumed
	float next[3] = { 0.f, 0.f, 0.f };
	if (active) {
#pragma unroll
		for (int c = 0; c < 3; c++) next[c] = __ldg(a.in + c * N + at);
	}
	float best = 0.0f;
	int buffer = 0;
#pragma unroll 1
	for (int r = r_first; r < r_stop; r++, at += Z) {
		const float v0 = next[0], v1 = next[1], v2 = next[2];
		if (active && r + 1 < r_load_end) {
#pragma unroll
			for (int c = 0; c < 3; c++) next[c] = __ldg(a.in + c * N + at + Z);
		} else {
			next[0] = next[1] = next[2] = 0.0f;
		}
		// axis-1 pass: row r is tap q of output row r + R - q (same chain as the axis-0 pass of stage 1)
#pragma unroll
		for (int q = K - 1; q >= 1; q--) {
			acc[0][q] = acc[0][q - 1] + v0 * a.k[q];
			acc[1][q] = acc[1][q - 1] + v1 * a.k[q];
			acc[2][q] = acc[2][q - 1] + v2 * a.k[q];
		}
		acc[0][0] = v0 * a.k[0];
		acc[1][0] = v1 * a.k[0];
		acc[2][0] = v2 * a.k[0];
		if (r - R < ys) continue;  // block-uniform: still priming
		float* row = row_memory + buffer * 3 * W;
		float w[3] = { 0.f, 0.f, 0.f };
		const int o = at - R * Z;  // voxel (x, r - R, z)
		if (active) {
			row[zl] = acc[0][K - 1];
			row[W + zl] = acc[1][K - 1];
			row[2 * W + zl] = acc[2][K - 1];
		}
		if (writes) {
			w[0] = a.warp[o];
			w[1] = a.warp[N + o];
			w[2] = a.warp[2 * N + o];
		}
		__syncthreads();
		if (writes) {
			// axis-2 pass
			float gq[3];
#pragma unroll
			for (int c = 0; c < 3; c++) {
				const float* v = row + c * W + tid;
				float sum = v[0] * a.k[0];
#pragma unroll
				for (int q = 1; q < K; q++) sum += v[q] * a.k[q];
				gq[c] = sum;
			}
			if (a.out != nullptr) {
				a.out[o] = gq[0];
				a.out[N + o] = gq[1];
				a.out[2 * N + o] = gq[2];
			}
			a.warp[o] = w[0] - gq[0] * a.rate;
			a.warp[N + o] = w[1] - gq[1] * a.rate;
			a.warp[2 * N + o] = w[2] - gq[2] * a.rate;
			float sq = gq[0] * gq[0];
			sq += gq[1] * gq[1];
			sq += gq[2] * gq[2];
			if (sq > best) best = sq;
		}
		buffer ^= 1;  // the row written two steps from now is read by nobody after the next barrier
	}
	if (a.max_sq_bits != nullptr) block_atomic_max(best, a.max_sq_bits + a.iteration);
}

// ---------------------------------------------------------------------------------------------- host-side launch helpers
struct TmaMaps {
	CUtensorMap g_prev, warp, canonical, pack;
	const void* pack_key = nullptr;
	const void* key[3] = { nullptr, nullptr, nullptr };  // pointers the maps were encoded for
	int tile_y = 0;                                       // 0: boxes of k_hier_stage1_tma; 1000 + TY: of k_hier_stage1_pair
};

inline bool tma_supported(const Grid3& g, const void* warp, const void* canonical, const void* g_prev) {
	auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
	return g.Z % 4 == 0 && g.N * 3 < (1ll << 31) && aligned(warp) && aligned(canonical) && aligned(g_prev);
}

template<bool TIKHONOV>
int ensure_maps(TmaMaps& maps, const Grid3& g, const float* warp, const float* canonical, const float* g_prev) {
	typedef Stage1Tile<TIKHONOV> T;
	if (maps.key[0] == warp && maps.key[1] == canonical && maps.key[2] == g_prev && maps.tile_y == 0) return LSF_OK;
	maps.tile_y = 0;
	LSF_TRY(make_planes_map(&maps.warp, warp, 3, g.N, g, T::TZ, T::TY));
	LSF_TRY(make_planes_map(&maps.canonical, canonical, 1, g.N, g, T::TZ, T::TY));
	LSF_TRY(make_planes_map(&maps.g_prev, g_prev, 3, g.N, g, T::GZ, T::GY));
	maps.key[0] = warp;
	maps.key[1] = canonical;
	maps.key[2] = g_prev;
	return LSF_OK;
}

template<typename T>
int ensure_pack_map(TmaMaps& maps, const HierIterArgs& a) {
	if (maps.pack_key == a.pack) return LSF_OK;
	LSF_TRY(make_pack_map(&maps.pack, a.pack, a.pack_X, a.g, T::PACK_BOX_Z, T::PACK_BOX_Y));
	maps.pack_key = a.pack;
	return LSF_OK;
}

// L2 prefetch of the pack by the tile leader (LSF_L2PF=0/1 overrides). Measured at 256^3: 4 % faster for the fused
// whole-iteration kernels (no Sobolev kernel), 3 % slower for stage 1 with the axis-0 pass -- on by default for the former.
inline bool l2_prefetch_enabled(bool fused_update) {
	const char* e = getenv("LSF_L2PF");
	if (e && (e[0] == '0' || e[0] == '1')) return e[0] == '1';
	return fused_update;
}

inline unsigned long long dup2_bits(float v) {
	unsigned bits;
	memcpy(&bits, &v, 4);
	return ((unsigned long long) bits << 32) | bits;
}

inline bool taps_are_symmetric(const Taps& taps) {
	for (int q = 0; q < taps.size; q++)
		if (memcmp(&taps.k[q], &taps.k[taps.size - 1 - q], sizeof(float)) != 0) return false;
	const char* e = getenv("LSF_SYM");  // A/B: LSF_SYM=0 multiplies every tap
	return !(e && e[0] == '0');
}

template<bool TIKHONOV, int R, bool DEC = true, bool APPLY = false, bool SYM = false>
int launch_stage1_tma(TmaMaps& maps, HierIterArgs a, const Taps& taps, float* h, int x_chunk, cudaStream_t stream) {
	typedef Stage1Tile<TIKHONOV> T;
	constexpr int NS = 4;
	if (APPLY && !SYM && taps_are_symmetric(taps))
		return launch_stage1_tma<TIKHONOV, R, DEC, APPLY, APPLY>(maps, a, taps, h, x_chunk, stream);
	LSF_TRY(ensure_maps<TIKHONOV>(maps, a.g, a.warp, a.canonical, a.g_prev));
	LSF_TRY(ensure_pack_map<T>(maps, a));
	XPassArgs t;
	for (int q = 0; q < 7; q++) {
		t.k[q] = q < 2 * R + 1 ? taps.k[q] : 0.0f;
		t.k2[q] = dup2_bits(t.k[q]);
	}
	t.x_chunk = x_chunk;
	t.one2 = F32X2_ONE;
	t.chunk_count = 0;
	a.g_out = h;
	const int pairs = a.batch_X > 0 ? a.g.X / a.batch_X : 1;
	a.batch_chunks = div_up(a.batch_X > 0 ? a.batch_X : a.g.X, x_chunk);
	const bool exact = APPLY && SYM && a.batch_X == 0 && a.g.Y % T::TY == 0 && a.g.Z % T::TZ == 0 && a.batch_chunks <= 12
			&& a.g.X < 32768;
	if (exact) {
		for (int c = 0; c <= a.batch_chunks; c++) t.chunk_bounds[c] = (short) std::min(c * x_chunk, a.g.X);
		t.chunk_count = a.batch_chunks;
		// Long-first schedule. When one plane holds between half a wave and a whole wave of tiles (256^3 on B200: 256 tiles,
		// 444 resident blocks) equal chunks end in a wave of equally long blocks on part of the SMs. Chunks of 55 % and 25 % of
		// the planes followed by three short ones let the short blocks fill the machine while the long ones finish:
		// measured 0.3432 - 0.3444 ms per 256^3 iteration against 0.3473 for five equal chunks (tools/xchunk_sweep.py,
		// profiles/r2_experiments.md; the slot model of marching_chunk does not predict this, the sweep does).
		// LSF_XSCHEDULE=0 keeps equal chunks.
		{
			static int sm_count = 0;
			if (sm_count == 0) {
				int device = 0;
				cudaGetDevice(&device);
				if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sm_count <= 0)
					sm_count = 148;
			}
			const char* off = getenv("LSF_XSCHEDULE");
			const int tiles = div_up(a.g.Z, T::TZ) * div_up(a.g.Y, T::TY), slots = 3 * sm_count;
			const int first = a.g.X * 142 / 256, second = a.g.X * 63 / 256, piece = (a.g.X - first - second) / 3;
			if (!(off && off[0] == '0') && tiles <= slots && slots < 2 * tiles && piece >= 2 * R + 2) {
				t.chunk_bounds[0] = 0;
				t.chunk_bounds[1] = (short) first;
				t.chunk_bounds[2] = (short) (first + second);
				t.chunk_bounds[3] = (short) (a.g.X - 2 * piece);
				t.chunk_bounds[4] = (short) (a.g.X - piece);
				t.chunk_bounds[5] = (short) a.g.X;
				t.chunk_count = 5;
				a.batch_chunks = 5;
			}
		}
		// experiment: LSF_XCHUNKS="142:68:22:12:12" (plane counts, must add up to X)
		const char* e = getenv("LSF_XCHUNKS");
		if (e) {
			int n = 0, at = 0;
			short bounds[13] = { 0 };
			const char* q = e;
			while (*q && n < 12) {
				at += atoi(q);
				bounds[++n] = (short) at;
				while (*q && *q != ',' && *q != ':') q++;
				if (*q) q++;
			}
			if (at == a.g.X) {
				for (int c = 0; c <= n; c++) t.chunk_bounds[c] = bounds[c];
				t.chunk_count = n;
				a.batch_chunks = n;
			}
		}
	}
	const dim3 block(T::TZ, T::TY, 1), grid(div_up(a.g.Z, T::TZ), div_up(a.g.Y, T::TY), pairs * a.batch_chunks);
	const size_t shared = (size_t) NS * T::STAGE_BYTES;
	static bool configured = false;
	if (!configured) {
		LSF_CUDA(cudaFuncSetAttribute(k_hier_stage1_tma<TIKHONOV, R, NS, DEC, false, 0, APPLY, false, SYM>,
				cudaFuncAttributeMaxDynamicSharedMemorySize, (int) shared));
		LSF_CUDA(cudaFuncSetAttribute(k_hier_stage1_tma<TIKHONOV, R, NS, DEC, false, 2, APPLY, false, SYM>,
				cudaFuncAttributeMaxDynamicSharedMemorySize, (int) shared));
		configured = true;
	}
	if (DEC && a.batch_X == 0 && l2_prefetch_enabled(false))
		LSF_CUDA(launch_dependent(k_hier_stage1_tma<TIKHONOV, R, NS, DEC, false, 2, APPLY, false, SYM>, grid, block, shared, stream,
				maps.g_prev, maps.warp, maps.canonical, maps.pack, a, t));
	else if (exact) {
		static bool configured_exact = false;
		if (!configured_exact) {
			LSF_CUDA(cudaFuncSetAttribute(k_hier_stage1_tma<TIKHONOV, R, NS, DEC, false, 0, APPLY, false, SYM, APPLY && SYM>,
					cudaFuncAttributeMaxDynamicSharedMemorySize, (int) shared));
			configured_exact = true;
		}
		LSF_CUDA(launch_dependent(k_hier_stage1_tma<TIKHONOV, R, NS, DEC, false, 0, APPLY, false, SYM, APPLY && SYM>, grid, block,
				shared, stream, maps.g_prev, maps.warp, maps.canonical, maps.pack, a, t));
	} else
		LSF_CUDA(launch_dependent(k_hier_stage1_tma<TIKHONOV, R, NS, DEC, false, 0, APPLY, false, SYM>, grid, block, shared, stream,
				maps.g_prev, maps.warp, maps.canonical, maps.pack, a, t));
	return LSF_OK;
}

// Whole iteration when no Sobolev kernel is configured: stage 1 + warp update + max-norm in one TMA-fed kernel.
// g_out (planes, may be nullptr without the Tikhonov term) must not alias a.g_prev; a.warp_out may alias a.warp.
template<bool TIKHONOV, bool SLAB = false>
int launch_stage1_fused_update(TmaMaps& maps, HierIterArgs a, int x_chunk, cudaStream_t stream) {
	typedef Stage1Tile<TIKHONOV> T;
	constexpr int NS = 4;
	LSF_TRY(ensure_maps<TIKHONOV>(maps, a.g, a.warp, a.canonical, a.g_prev));
	XPassArgs t;
	for (int q = 0; q < 7; q++) {
		t.k[q] = 0.0f;
		t.k2[q] = 0ull;
	}
	t.x_chunk = x_chunk;
	t.one2 = F32X2_ONE;
	t.chunk_count = 0;
	const int pairs = (!SLAB && a.batch_X > 0) ? a.g.X / a.batch_X : 1;
	a.batch_chunks = div_up(pairs > 1 || a.batch_X > 0 ? a.batch_X : a.x_end - a.x_begin, x_chunk);
	const dim3 block(T::TZ, T::TY, 1), grid(div_up(a.g.Z, T::TZ), div_up(a.g.Y, T::TY), pairs * a.batch_chunks);
	const size_t shared = (size_t) NS * T::STAGE_BYTES;
	LSF_TRY(ensure_pack_map<T>(maps, a));
	static bool configured = false;
	if (!configured) {
		LSF_CUDA(cudaFuncSetAttribute(k_hier_stage1_tma<TIKHONOV, 0, NS, true, true, 0, false, SLAB>,
				cudaFuncAttributeMaxDynamicSharedMemorySize, (int) shared));
		LSF_CUDA(cudaFuncSetAttribute(k_hier_stage1_tma<TIKHONOV, 0, NS, true, true, 2, false, SLAB>,
				cudaFuncAttributeMaxDynamicSharedMemorySize, (int) shared));
		configured = true;
	}
	if (!SLAB && a.batch_X == 0 && l2_prefetch_enabled(true))
		k_hier_stage1_tma<TIKHONOV, 0, NS, true, true, 2, false, SLAB> <<<counted(grid), block, shared, stream>>>(maps.g_prev, maps.warp,
				maps.canonical, maps.pack, a, t);
	else
		k_hier_stage1_tma<TIKHONOV, 0, NS, true, true, 0, false, SLAB> <<<counted(grid), block, shared, stream>>>(maps.g_prev, maps.warp,
				maps.canonical, maps.pack, a, t);
	return LSF_OK;
}

template<int R>
void launch_ymarch(const Taps& taps, const HierIterArgs& a, const float* h, float* filtered, float* warp, int y_chunk,
		cudaStream_t stream) {
	const Grid3& g = a.g;
	YMarchArgs f;
	f.in = h;
	f.out = filtered;
	f.warp = warp;
	f.g = g;
	for (int q = 0; q < 7; q++) f.k[q] = q < 2 * R + 1 ? taps.k[q] : 0.0f;
	f.rate = a.rate;
	f.threshold = a.threshold;
	f.max_sq_bits = a.max_sq_bits;
	f.iteration = a.iteration;
	f.check_convergence = a.check_convergence;
	f.y_chunk = y_chunk;
	f.x_begin = 0;
	f.tile_z = std::min(256, (int) div_up(g.Z, 32) * 32);
	const int tiles = div_up(g.Z, f.tile_z);
	const dim3 grid(tiles, g.X, div_up(g.Y, y_chunk));
	const int threads = f.tile_z + (tiles > 1 ? 32 : 0);
	const size_t shared = (size_t) 6 * (f.tile_z + 2 * R) * sizeof(float);
	k_sobolev_ymarch<R> <<<counted(grid), threads, shared, stream>>>(f);
}

// One iteration = stage 1 with the axis-0 pass, then the y-marching filter. Returns a negative status on failure.
template<int R>
int launch_tma_iteration(bool tikhonov, TmaMaps& maps, HierIterArgs a, const Taps& taps, float* h, float* filtered,
		float* warp, int x_chunk, int y_chunk, cudaStream_t stream, cudaEvent_t* events) {
	const bool coupled = getenv("LSF_DECOUPLE") && getenv("LSF_DECOUPLE")[0] == '0';  // A/B: block barrier per plane
	if (tikhonov && coupled) LSF_TRY((launch_stage1_tma<true, R, false>(maps, a, taps, h, x_chunk, stream)));
	else if (tikhonov) LSF_TRY((launch_stage1_tma<true, R>(maps, a, taps, h, x_chunk, stream)));
	else LSF_TRY((launch_stage1_tma<false, R>(maps, a, taps, h, x_chunk, stream)));
	if (events) cudaEventRecord(events[1], stream);
	launch_ymarch<R>(taps, a, h, filtered, warp, y_chunk, stream);
	if (events) cudaEventRecord(events[2], stream);
	return LSF_OK;
}

#endif  // __CUDACC__

}  // namespace lsf
