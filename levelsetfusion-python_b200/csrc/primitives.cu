// primitives.cu -- C-ABI entry points of the field-math primitives (reference cpp/src/math/*,
// cpp/src/nonrigid_optimization/field_warping.*). They run the same device code the optimizers use
// (gather4 on the padded pack, laplace_term, k_convolve_axis*, restrict / prolong kernels) on API-layout
// (interleaved) arrays, so the parity tests of the primitives exercise the optimizer's kernels.
#include "kernels2d.cuh"
#include "kernels3d_fused.cuh"

#include <cstdlib>

namespace lsf {
namespace {

inline unsigned blocks_for(long long n) {
	return div_up(n, 256);
}

struct Staged {
	Arena arena;
	cudaStream_t stream;
	int memory_kind;
	explicit Staged(void* stream_handle, int kind) :
			arena(static_cast<cudaStream_t>(stream_handle)), stream(static_cast<cudaStream_t>(stream_handle)),
					memory_kind(kind) {
	}
	int in(const float* src, size_t count, const float** out) {
		return to_device(arena, src, count, memory_kind, stream, out);
	}
	int out_buffer(float* user, size_t count, float** dev) {
		if (memory_kind == LSF_DEVICE) {
			*dev = user;
			return LSF_OK;
		}
		return arena.alloc(dev, count);
	}
	int finish(const float* dev, float* user, size_t count) {
		LSF_CUDA(cudaGetLastError());
		return from_device(dev, user, count, memory_kind, stream);
	}
};

}  // namespace
}  // namespace lsf

using namespace lsf;

// ------------------------------------------------------------------------------------------------ gather
extern "C" int lsf_warp_3d(const float* field, int channels, const float* warp, int X, int Y, int Z, float oob_value,
		float* out, int memory_kind, void* stream_handle) {
	LSF_REQUIRE(channels == 1 || channels == 3, "3D fields have 1 or 3 channels, got %d", channels);
	LSF_REQUIRE(X > 0 && Y > 0 && Z > 0 && field && warp && out, "invalid arguments");
	Staged st(stream_handle, memory_kind);
	const Grid3 g(X, Y, Z);
	const float *field_dev, *warp_dev;
	float* out_dev;
	LSF_TRY(st.in(field, (size_t) g.N * channels, &field_dev));
	LSF_TRY(st.in(warp, (size_t) g.N * 3, &warp_dev));
	LSF_TRY(st.out_buffer(out, (size_t) g.N * channels, &out_dev));
	float4* pack;
	LSF_TRY(st.arena.alloc(&pack, (size_t) g.padded_count()));
	k_fill4<<<counted(blocks_for(g.padded_count())), 256, 0, st.stream>>>(pack, g.padded_count(),
			make_float4(oob_value, oob_value, oob_value, oob_value));
	k_pack_field3d<<<counted(grid3(g)), block3(), 0, st.stream>>>(field_dev, channels, pack, g);
	k_gather_pack3d<<<counted(grid3(g)), block3(), 0, st.stream>>>(pack, warp_dev, out_dev, channels, g);
	return st.finish(out_dev, out, (size_t) g.N * channels);
}

extern "C" int lsf_warp_2d(const float* field, int channels, const float* warp, int H, int W, float oob_value,
		float* out, int memory_kind, void* stream_handle) {
	LSF_REQUIRE(channels == 1 || channels == 2, "2D fields have 1 or 2 channels, got %d", channels);
	LSF_REQUIRE(H > 0 && W > 0 && field && warp && out, "invalid arguments");
	Staged st(stream_handle, memory_kind);
	const Grid2 g(H, W);
	const float *field_dev, *warp_dev;
	float* out_dev;
	LSF_TRY(st.in(field, (size_t) g.N * channels, &field_dev));
	LSF_TRY(st.in(warp, (size_t) g.N * 2, &warp_dev));
	LSF_TRY(st.out_buffer(out, (size_t) g.N * channels, &out_dev));
	float4* pack;
	LSF_TRY(st.arena.alloc(&pack, (size_t) g.padded_count()));
	k_fill4<<<counted(blocks_for(g.padded_count())), 256, 0, st.stream>>>(pack, g.padded_count(),
			make_float4(oob_value, oob_value, oob_value, oob_value));
	k_pack_field2d<<<counted(grid2(g)), block3(), 0, st.stream>>>(field_dev, channels, pack, g);
	k_gather_pack2d<<<counted(grid2(g)), block3(), 0, st.stream>>>(pack, warp_dev, out_dev, channels, g);
	return st.finish(out_dev, out, (size_t) g.N * channels);
}

// ------------------------------------------------------------------------------------------------ gradient
extern "C" int lsf_gradient_3d(const float* field, int X, int Y, int Z, float* out, int memory_kind,
		void* stream_handle) {
	LSF_REQUIRE(X > 0 && Y > 0 && Z > 0 && field && out, "invalid arguments");
	Staged st(stream_handle, memory_kind);
	const Grid3 g(X, Y, Z);
	const float* field_dev;
	float* out_dev;
	LSF_TRY(st.in(field, (size_t) g.N, &field_dev));
	LSF_TRY(st.out_buffer(out, (size_t) g.N * 3, &out_dev));
	k_gradient3d<<<counted(grid3(g)), block3(), 0, st.stream>>>(field_dev, out_dev, g);
	return st.finish(out_dev, out, (size_t) g.N * 3);
}

extern "C" int lsf_gradient_2d(const float* field, int H, int W, float* out, int memory_kind, void* stream_handle) {
	LSF_REQUIRE(H > 0 && W > 0 && field && out, "invalid arguments");
	Staged st(stream_handle, memory_kind);
	const Grid2 g(H, W);
	const float* field_dev;
	float* out_dev;
	LSF_TRY(st.in(field, (size_t) g.N, &field_dev));
	LSF_TRY(st.out_buffer(out, (size_t) g.N * 2, &out_dev));
	k_gradient2d<<<counted(grid2(g)), block3(), 0, st.stream>>>(field_dev, out_dev, g);
	return st.finish(out_dev, out, (size_t) g.N * 2);
}

// ------------------------------------------------------------------------------------------------ laplacian
extern "C" int lsf_laplacian_3d(const float* vfield, int X, int Y, int Z, float* out, int memory_kind,
		void* stream_handle) {
	LSF_REQUIRE(X > 0 && Y > 0 && Z > 0 && vfield && out, "invalid arguments");
	Staged st(stream_handle, memory_kind);
	const Grid3 g(X, Y, Z);
	const float* in_dev;
	float *out_dev, *planes_in, *planes_out;
	LSF_TRY(st.in(vfield, (size_t) g.N * 3, &in_dev));
	LSF_TRY(st.out_buffer(out, (size_t) g.N * 3, &out_dev));
	LSF_TRY(st.arena.alloc(&planes_in, (size_t) g.N * 3));
	LSF_TRY(st.arena.alloc(&planes_out, (size_t) g.N * 3));
	k_aos_to_planes<<<counted(blocks_for(g.N)), 256, 0, st.stream>>>(in_dev, planes_in, g.N, 3);
	k_laplacian_planes3d<<<counted(grid3(g)), block3(), 0, st.stream>>>(planes_in, planes_out, 3, g);
	k_planes_to_aos<<<counted(blocks_for(g.N)), 256, 0, st.stream>>>(planes_out, out_dev, g.N, 3);
	return st.finish(out_dev, out, (size_t) g.N * 3);
}

extern "C" int lsf_laplacian_2d(const float* vfield, int H, int W, float* out, int memory_kind, void* stream_handle) {
	LSF_REQUIRE(H > 0 && W > 0 && vfield && out, "invalid arguments");
	Staged st(stream_handle, memory_kind);
	const Grid2 g(H, W);
	const float* in_dev;
	float *out_dev, *planes_in, *planes_out;
	LSF_TRY(st.in(vfield, (size_t) g.N * 2, &in_dev));
	LSF_TRY(st.out_buffer(out, (size_t) g.N * 2, &out_dev));
	LSF_TRY(st.arena.alloc(&planes_in, (size_t) g.N * 2));
	LSF_TRY(st.arena.alloc(&planes_out, (size_t) g.N * 2));
	k_aos_to_planes<<<counted(blocks_for(g.N)), 256, 0, st.stream>>>(in_dev, planes_in, g.N, 2);
	k_laplacian_planes2d<<<counted(grid2(g)), block3(), 0, st.stream>>>(planes_in, planes_out, 2, g);
	k_planes_to_aos<<<counted(blocks_for(g.N)), 256, 0, st.stream>>>(planes_out, out_dev, g.N, 2);
	return st.finish(out_dev, out, (size_t) g.N * 2);
}

// ------------------------------------------------------------------------------------------------ separable filter
extern "C" int lsf_convolve_3d(float* vfield, int X, int Y, int Z, const float* kernel, int kernel_size,
		int memory_kind, void* stream_handle) {
	LSF_REQUIRE(X > 0 && Y > 0 && Z > 0 && vfield, "invalid arguments");
	Taps taps;
	LSF_TRY(make_taps(kernel, kernel_size, &taps));
	Staged st(stream_handle, memory_kind);
	const Grid3 g(X, Y, Z);
	const float* in_dev;
	float *out_dev, *a, *b;
	LSF_TRY(st.in(vfield, (size_t) g.N * 3, &in_dev));
	LSF_TRY(st.out_buffer(vfield, (size_t) g.N * 3, &out_dev));
	LSF_TRY(st.arena.alloc(&a, (size_t) g.N * 3));
	LSF_TRY(st.arena.alloc(&b, (size_t) g.N * 3));
	k_aos_to_planes<<<counted(blocks_for(g.N)), 256, 0, st.stream>>>(in_dev, a, g.N, 3);
	const char* legacy = getenv("LSF_LEGACY_KERNELS");
	if (!(legacy && legacy[0] == '1')
			&& launch_fused_filter_any(taps, 0.0f, 0.0f, g, a, b, nullptr, nullptr, 0, 0, st.stream)) {
		// the optimizer's fused three-pass kernel, filter only
		k_planes_to_aos<<<counted(blocks_for(g.N)), 256, 0, st.stream>>>(b, out_dev, g.N, 3);
		return st.finish(out_dev, vfield, (size_t) g.N * 3);
	}
	ConvArgs c;
	c.g = g;
	c.taps = taps;
	c.rate = 0.0f;
	c.threshold = 0.0f;
	c.max_sq_bits = nullptr;
	c.iteration = 0;
	c.check_convergence = 0;
	c.channels = 3;
	c.warp = nullptr;
	c.in = a;
	c.out = b;
	k_convolve_axis3d<0, false> <<<counted(grid3(g)), block3(), 0, st.stream>>>(c);
	c.in = b;
	c.out = a;
	k_convolve_axis3d<1, false> <<<counted(grid3(g)), block3(), 0, st.stream>>>(c);
	c.in = a;
	c.out = b;
	k_convolve_axis3d<2, false> <<<counted(grid3(g)), block3(), 0, st.stream>>>(c);
	k_planes_to_aos<<<counted(blocks_for(g.N)), 256, 0, st.stream>>>(b, out_dev, g.N, 3);
	return st.finish(out_dev, vfield, (size_t) g.N * 3);
}

extern "C" int lsf_convolve_2d(float* vfield, int H, int W, const float* kernel, int kernel_size, int preserve_zeros,
		int memory_kind, void* stream_handle) {
	LSF_REQUIRE(H > 0 && W > 0 && vfield, "invalid arguments");
	Taps taps;
	LSF_TRY(make_taps(kernel, kernel_size, &taps));
	Staged st(stream_handle, memory_kind);
	const Grid2 g(H, W);
	const float* in_dev;
	float *out_dev, *a, *b;
	LSF_TRY(st.in(vfield, (size_t) g.N * 2, &in_dev));
	LSF_TRY(st.out_buffer(vfield, (size_t) g.N * 2, &out_dev));
	LSF_TRY(st.arena.alloc(&a, (size_t) g.N * 2));
	LSF_TRY(st.arena.alloc(&b, (size_t) g.N * 2));
	k_aos_to_planes<<<counted(blocks_for(g.N)), 256, 0, st.stream>>>(in_dev, a, g.N, 2);
	ConvArgs2 c;
	c.g = g;
	c.taps = taps;
	c.rate = 0.0f;
	c.threshold = 0.0f;
	c.max_sq_bits = nullptr;
	c.iteration = 0;
	c.check_convergence = 0;
	c.channels = 2;
	c.preserve_zeros = preserve_zeros ? 1 : 0;
	c.warp = nullptr;
	c.in = a;
	c.out = b;
	k_convolve_axis2d<0, false> <<<counted(grid2(g)), block3(), 0, st.stream>>>(c);
	c.in = b;
	c.out = a;
	k_convolve_axis2d<1, false> <<<counted(grid2(g)), block3(), 0, st.stream>>>(c);
	k_planes_to_aos<<<counted(blocks_for(g.N)), 256, 0, st.stream>>>(a, out_dev, g.N, 2);
	return st.finish(out_dev, vfield, (size_t) g.N * 2);
}

// ------------------------------------------------------------------------------------------------ restrict / prolong
extern "C" int lsf_downsample_3d(const float* field, int channels, int X, int Y, int Z, int linear, float* out,
		int memory_kind, void* stream_handle) {
	LSF_REQUIRE(channels == 1 || channels == 3, "3D fields have 1 or 3 channels, got %d", channels);
	LSF_REQUIRE(X > 1 && Y > 1 && Z > 1 && field && out, "invalid arguments");
	if (linear)
		LSF_REQUIRE(X % 2 == 0 && Y % 2 == 0 && Z % 2 == 0 && X > 2 && Y > 2 && Z > 2,
				"Each dimension of the argument 'field' must be divisible by 2 and greater than 2.");
	Staged st(stream_handle, memory_kind);
	const Grid3 g(X, Y, Z), d = g.half();
	const float* in_dev;
	float *out_dev, *planes_in, *planes_out;
	LSF_TRY(st.in(field, (size_t) g.N * channels, &in_dev));
	LSF_TRY(st.out_buffer(out, (size_t) d.N * channels, &out_dev));
	LSF_TRY(st.arena.alloc(&planes_in, (size_t) g.N * channels));
	LSF_TRY(st.arena.alloc(&planes_out, (size_t) d.N * channels));
	k_aos_to_planes<<<counted(blocks_for(g.N)), 256, 0, st.stream>>>(in_dev, planes_in, g.N, channels);
	for (int c = 0; c < channels; c++) {
		PlainAccess access { planes_in + c * g.N, planes_out + c * d.N };
		if (linear) k_downsample_linear3d<PlainAccess> <<<counted(grid3(d)), block3(), 0, st.stream>>>(access, g, d);
		else k_downsample_average3d<PlainAccess> <<<counted(grid3(d)), block3(), 0, st.stream>>>(access, g, d);
	}
	k_planes_to_aos<<<counted(blocks_for(d.N)), 256, 0, st.stream>>>(planes_out, out_dev, d.N, channels);
	return st.finish(out_dev, out, (size_t) d.N * channels);
}

extern "C" int lsf_upsample_3d(const float* field, int channels, int X, int Y, int Z, int linear, float* out,
		int memory_kind, void* stream_handle) {
	LSF_REQUIRE(channels == 1 || channels == 3, "3D fields have 1 or 3 channels, got %d", channels);
	LSF_REQUIRE(X > 0 && Y > 0 && Z > 0 && field && out, "invalid arguments");
	if (linear) LSF_REQUIRE(X > 1 && Y > 1 && Z > 1, "linear upsampling needs at least 2 elements per dimension");
	Staged st(stream_handle, memory_kind);
	const Grid3 g(X, Y, Z), d = g.twice();
	const float* in_dev;
	float *out_dev, *planes_in, *planes_out;
	LSF_TRY(st.in(field, (size_t) g.N * channels, &in_dev));
	LSF_TRY(st.out_buffer(out, (size_t) d.N * channels, &out_dev));
	LSF_TRY(st.arena.alloc(&planes_in, (size_t) g.N * channels));
	LSF_TRY(st.arena.alloc(&planes_out, (size_t) d.N * channels));
	k_aos_to_planes<<<counted(blocks_for(g.N)), 256, 0, st.stream>>>(in_dev, planes_in, g.N, channels);
	if (linear) k_upsample_linear3d<<<counted(grid3(d)), block3(), 0, st.stream>>>(planes_in, planes_out, channels, g, d);
	else k_upsample_nearest3d<<<counted(grid3(d)), block3(), 0, st.stream>>>(planes_in, planes_out, channels, g, d);
	k_planes_to_aos<<<counted(blocks_for(d.N)), 256, 0, st.stream>>>(planes_out, out_dev, d.N, channels);
	return st.finish(out_dev, out, (size_t) d.N * channels);
}

extern "C" int lsf_downsample_2d(const float* field, int channels, int H, int W, int linear, float* out,
		int memory_kind, void* stream_handle) {
	LSF_REQUIRE(channels == 1 || channels == 2, "2D fields have 1 or 2 channels, got %d", channels);
	LSF_REQUIRE(H > 1 && W > 1 && field && out, "invalid arguments");
	if (linear)
		LSF_REQUIRE(H % 2 == 0 && W % 2 == 0 && H > 2 && W > 2,
				"Each dimension of the argument 'field' must be divisible by 2 and greater than 2.");
	else
		LSF_REQUIRE(is_power_of_two(H) && is_power_of_two(W),
				"The argument 'field' must have a power of two for each dimension.");
	Staged st(stream_handle, memory_kind);
	const Grid2 g(H, W), d = g.half();
	const float* in_dev;
	float *out_dev, *planes_in, *planes_out;
	LSF_TRY(st.in(field, (size_t) g.N * channels, &in_dev));
	LSF_TRY(st.out_buffer(out, (size_t) d.N * channels, &out_dev));
	LSF_TRY(st.arena.alloc(&planes_in, (size_t) g.N * channels));
	LSF_TRY(st.arena.alloc(&planes_out, (size_t) d.N * channels));
	k_aos_to_planes<<<counted(blocks_for(g.N)), 256, 0, st.stream>>>(in_dev, planes_in, g.N, channels);
	for (int c = 0; c < channels; c++) {
		PlainAccess2 access { planes_in + c * g.N, planes_out + c * d.N };
		if (linear) k_downsample_linear2d<PlainAccess2> <<<counted(grid2(d)), block3(), 0, st.stream>>>(access, g, d);
		else k_downsample_average2d<PlainAccess2> <<<counted(grid2(d)), block3(), 0, st.stream>>>(access, g, d);
	}
	k_planes_to_aos<<<counted(blocks_for(d.N)), 256, 0, st.stream>>>(planes_out, out_dev, d.N, channels);
	return st.finish(out_dev, out, (size_t) d.N * channels);
}

extern "C" int lsf_upsample_2d(const float* field, int channels, int H, int W, int linear, float* out, int memory_kind,
		void* stream_handle) {
	LSF_REQUIRE(channels == 1 || channels == 2, "2D fields have 1 or 2 channels, got %d", channels);
	LSF_REQUIRE(H > 0 && W > 0 && field && out, "invalid arguments");
	if (linear) LSF_REQUIRE(H > 1 && W > 1, "linear upsampling needs at least 2 elements per dimension");
	Staged st(stream_handle, memory_kind);
	const Grid2 g(H, W), d(2 * H, 2 * W);
	const float* in_dev;
	float *out_dev, *planes_in, *planes_out;
	LSF_TRY(st.in(field, (size_t) g.N * channels, &in_dev));
	LSF_TRY(st.out_buffer(out, (size_t) d.N * channels, &out_dev));
	LSF_TRY(st.arena.alloc(&planes_in, (size_t) g.N * channels));
	LSF_TRY(st.arena.alloc(&planes_out, (size_t) d.N * channels));
	k_aos_to_planes<<<counted(blocks_for(g.N)), 256, 0, st.stream>>>(in_dev, planes_in, g.N, channels);
	k_upsample2d<<<counted(grid2(d)), block3(), 0, st.stream>>>(planes_in, planes_out, channels, g, d, linear ? 1 : 0);
	k_planes_to_aos<<<counted(blocks_for(d.N)), 256, 0, st.stream>>>(planes_out, out_dev, d.N, channels);
	return st.finish(out_dev, out, (size_t) d.N * channels);
}

// ------------------------------------------------------------------------------------------------ max norm
extern "C" int lsf_max_norm(const float* vfield, int channels, long long count, float* max_norm_out, int memory_kind,
		void* stream_handle) {
	LSF_REQUIRE(channels > 0 && count > 0 && vfield && max_norm_out, "invalid arguments");
	Staged st(stream_handle, memory_kind);
	const float* in_dev;
	float* planes;
	unsigned* bits;
	LSF_TRY(st.in(vfield, (size_t) count * channels, &in_dev));
	LSF_TRY(st.arena.alloc(&planes, (size_t) count * channels));
	LSF_TRY(st.arena.alloc(&bits, 1));
	LSF_CUDA(cudaMemsetAsync(bits, 0, sizeof(unsigned), st.stream));
	k_aos_to_planes<<<counted(blocks_for(count)), 256, 0, st.stream>>>(in_dev, planes, count, channels);
	const unsigned blocks = (unsigned) std::min<long long>(blocks_for(count), 148 * 8);
	k_max_sq_norm_planes<<<counted(blocks), 256, 0, st.stream>>>(planes, count, channels, bits);
	LSF_CUDA(cudaGetLastError());
	unsigned host_bits = 0;
	LSF_CUDA(cudaMemcpyAsync(&host_bits, bits, sizeof(unsigned), cudaMemcpyDeviceToHost, st.stream));
	LSF_CUDA(cudaStreamSynchronize(st.stream));
	float sq;
	memcpy(&sq, &host_bits, sizeof(float));
	*max_norm_out = sqrtf(sq);
	return LSF_OK;
}

// maximum vector length WITH its location (reference math::locate_max_norm, cpp/src/math/statistics.tpp:57-100): the key
// (bits of the squared length << 32 | ~order) makes one atomicMax return the maximum and, among equal maxima, the element
// the reference's traversal meets first (its strict `>` keeps the first one). order = index in the reference's
// column-major element order: 2D [H][W] field c * H + r, 3D [X][Y][Z] field x + X * (y + Y * z).
namespace lsf {
namespace {

__global__ void __launch_bounds__(256) k_locate_max_norm(const float* __restrict__ vfield, int channels, int nd, int d0, int d1,
		int d2, unsigned long long* __restrict__ best_key) {
	const long long n = (long long) d0 * d1 * d2;
	unsigned long long best = 0;
	for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
		float sq = 0.0f;
		for (int c = 0; c < channels; c++) sq += vfield[i * channels + c] * vfield[i * channels + c];
		if (!(sq >= 0.0f)) continue;  // NaN: never greater than the running maximum
		unsigned order;
		if (nd == 2) order = (unsigned) ((i % d1) * d0 + i / d1);
		else {
			const long long z = i % d2, y = (i / d2) % d1, x = i / ((long long) d1 * d2);
			order = (unsigned) (x + (long long) d0 * (y + (long long) d1 * z));
		}
		const unsigned long long key = ((unsigned long long) __float_as_uint(sq) << 32) | (0xffffffffu - order);
		best = key > best ? key : best;
	}
#pragma unroll
	for (int offset = 16; offset > 0; offset >>= 1) {
		const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, offset);
		best = other > best ? other : best;
	}
	if ((threadIdx.x & 31) == 0) atomicMax(best_key, best);
}

}  // namespace
}  // namespace lsf

extern "C" int lsf_locate_max_norm(const float* vfield, int channels, int nd, const int* dims, float* max_norm_out,
		int* coordinates_out, int memory_kind, void* stream_handle) {
	LSF_REQUIRE(channels > 0 && (nd == 2 || nd == 3) && dims && vfield && max_norm_out && coordinates_out, "invalid arguments");
	const int d0 = dims[0], d1 = dims[1], d2 = nd == 3 ? dims[2] : 1;
	LSF_REQUIRE(d0 > 0 && d1 > 0 && d2 > 0 && (long long) d0 * d1 * d2 < (1ll << 32), "invalid field dimensions");
	const long long count = (long long) d0 * d1 * d2;
	Staged st(stream_handle, memory_kind);
	const float* in_dev;
	unsigned long long* key;
	LSF_TRY(st.in(vfield, (size_t) count * channels, &in_dev));
	LSF_TRY(st.arena.alloc(&key, 1));
	LSF_CUDA(cudaMemsetAsync(key, 0, sizeof(unsigned long long), st.stream));
	const unsigned blocks = (unsigned) std::min<long long>(blocks_for(count), 148 * 8);
	k_locate_max_norm<<<counted(blocks), 256, 0, st.stream>>>(in_dev, channels, nd, d0, d1, d2, key);
	LSF_CUDA(cudaGetLastError());
	unsigned long long host_key = 0;
	LSF_CUDA(cudaMemcpyAsync(&host_key, key, sizeof(host_key), cudaMemcpyDeviceToHost, st.stream));
	LSF_CUDA(cudaStreamSynchronize(st.stream));
	const unsigned bits = (unsigned) (host_key >> 32);
	float sq;
	memcpy(&sq, &bits, sizeof(float));
	*max_norm_out = sqrtf(sq);
	const unsigned order = 0xffffffffu - (unsigned) (host_key & 0xffffffffull);
	if (nd == 2) {
		// statistics.tpp:70-71: x = i_element / column_count, y = i_element % column_count (= column, row of a square field)
		coordinates_out[0] = (int) (order / (unsigned) d1);
		coordinates_out[1] = (int) (order % (unsigned) d1);
		coordinates_out[2] = 0;
	} else {
		coordinates_out[0] = (int) (order % (unsigned) d0);
		coordinates_out[1] = (int) ((order / (unsigned) d0) % (unsigned) d1);
		coordinates_out[2] = (int) (order / ((unsigned) d0 * (unsigned) d1));
	}
	return LSF_OK;
}
