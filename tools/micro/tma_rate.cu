// Micro-benchmark: how fast does one SM's TMA unit deliver 4-D boxes with short rows?  (profiles/r2_experiments.md)
// One block per SM, one producer thread keeps a ring of STAGES boxes in flight (cp.async.bulk.tensor.4d, completion on an
// mbarrier, re-issued as soon as a box has landed; nobody reads the data). Boxes walk through a [C][X][Y][Z] float
// volume; `window_planes` limits the walk to the first planes (small window = L2-resident source).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../levelsetfusion-python_b200/csrc -o tma_rate tma_rate.cu
// Run:   ./tma_rate            (prints bytes / cycle / SM and cycles per box row for a list of box shapes)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "kernels3d_tma.cuh"

namespace lsf {
void set_error(const char* fmt, ...) { va_list a; va_start(a, fmt); vfprintf(stderr, fmt, a); va_end(a); fprintf(stderr, "\n"); }
void count_launches(int) {}
}
using namespace lsf;

constexpr int STAGES_MAX = 8;

__global__ void __launch_bounds__(64, 1) k_rate(const __grid_constant__ CUtensorMap map, int box_bytes, int stages, int boxes,
		int bx, int by, int bz, int X, int Y, int Z, int window_planes, long long* cycles, int align_z) {
	extern __shared__ __align__(128) unsigned char buf[];
	__shared__ uint64_t bar[STAGES_MAX];
	if (threadIdx.x == 0) {
		for (int s = 0; s < stages; s++) mbar_init(&bar[s], 1);
		mbar_fence_init();
	}
	__syncthreads();
	if (threadIdx.x != 0) return;
	const int nx = window_planes / bx, ny = Y / by, nz = Z / bz;
	const int per = nx * ny * nz;
	const long long t0 = clock64();
	for (int n = 0; n < boxes + stages; n++) {
		const int s = n % stages;
		if (n >= stages) mbar_wait(&bar[s], ((n / stages) - 1) & 1);
		if (n < boxes) {
			// a different box per block and step, spread over the window
			const unsigned k = ((unsigned) n * gridDim.x + blockIdx.x) * 2654435761u % (unsigned) per;
			const int z0 = (k % nz) * bz - align_z, y0 = ((k / nz) % ny) * by, x0 = (k / (nz * ny)) * bx;
			mbar_expect_tx(&bar[s], box_bytes);
			tma_load_4d(buf + (size_t) s * box_bytes, &map, z0, y0, x0, 0, &bar[s]);
		}
	}
	cycles[blockIdx.x] = clock64() - t0;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main() {
	const int X = 256, Y = 256, Z = 256, C = 3;
	float* dev;
	CK(cudaMalloc(&dev, (size_t) C * X * Y * Z * 4));
	CK(cudaMemset(dev, 0, (size_t) C * X * Y * Z * 4));
	long long* cycles;
	CK(cudaMalloc(&cycles, 148 * 8));
	EncodeTiledFn encode = encode_tiled_fn();
	if (!encode) { printf("no encoder\n"); return 1; }
	struct Shape { int bz, by, bx, c, align; };
	const Shape shapes[] = { { 32, 8, 8, 3, 0 }, { 32, 8, 14, 3, 0 }, { 40, 10, 10, 3, 4 }, { 40, 8, 8, 3, 4 }, { 64, 8, 8, 3, 0 },
			{ 64, 8, 4, 3, 0 }, { 128, 8, 2, 3, 0 }, { 256, 4, 2, 3, 0 }, { 32, 8, 1, 3, 0 }, { 32, 10, 1, 3, 0 }, { 64, 10, 1, 3, 0 },
			{ 40, 10, 1, 3, 4 }, { 32, 32, 1, 3, 0 }, { 16, 16, 8, 3, 0 } };
	printf("%-22s %7s %6s %7s %9s %12s %12s %12s\n", "box z x y x x x c", "bytes", "rows", "stages", "window", "B/cyc/SM", "cyc/row", "GB/s chip");
	for (const Shape& sh : shapes)
		for (int window : { 8, 256 })
			for (int stages : { 2, 4 }) {
				const cuuint64_t dims[4] = { (cuuint64_t) Z, (cuuint64_t) Y, (cuuint64_t) X, (cuuint64_t) C };
				const cuuint64_t strides[3] = { (cuuint64_t) Z * 4, (cuuint64_t) Y * Z * 4, (cuuint64_t) X * Y * Z * 4 };
				const cuuint32_t box[4] = { (cuuint32_t) sh.bz, (cuuint32_t) sh.by, (cuuint32_t) sh.bx, (cuuint32_t) sh.c };
				const cuuint32_t es[4] = { 1, 1, 1, 1 };
				CUtensorMap map;
				if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dev, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
						CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
					printf("encode failed\n");
					continue;
				}
				const int box_bytes = sh.bz * sh.by * sh.bx * sh.c * 4, rows = sh.by * sh.bx * sh.c;
				const int smem = box_bytes * stages;
				if (smem > 200 * 1024 || (window < sh.bx)) continue;
				CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
				const int boxes = (int) (48ll * 1024 * 1024 / box_bytes / 16);
				float best = 1e30f;
				long long best_cycles = 0;
				for (int rep = 0; rep < 3; rep++) {
					cudaEvent_t a, b;
					cudaEventCreate(&a); cudaEventCreate(&b);
					cudaEventRecord(a);
					k_rate<<<148, 64, smem>>>(map, box_bytes, stages, boxes, sh.bx, sh.by, sh.bz, X, Y, Z, window, cycles, sh.align);
					cudaEventRecord(b);
					CK(cudaDeviceSynchronize());
					float ms; cudaEventElapsedTime(&ms, a, b);
					std::vector<long long> h(148);
					CK(cudaMemcpy(h.data(), cycles, 148 * 8, cudaMemcpyDeviceToHost));
					long long mx = 0; for (long long v : h) mx = v > mx ? v : mx;
					if (ms < best) { best = ms; best_cycles = mx; }
				}
				const double bytes_per_sm = (double) boxes * box_bytes;
				char name[64];
				snprintf(name, sizeof name, "%d x %d x %d x %d", sh.bz, sh.by, sh.bx, sh.c);
				printf("%-22s %7d %6d %7d %9d %12.2f %12.2f %12.1f\n", name, box_bytes, rows, stages, window, bytes_per_sm / best_cycles,
						(double) best_cycles / ((double) boxes * rows), 148.0 * bytes_per_sm / (best * 1e-3) / 1e9);
			}
	return 0;
}
