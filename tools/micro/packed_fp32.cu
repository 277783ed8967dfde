// Microbenchmark: issue rate of scalar FMUL/FADD vs packed FMUL2/FADD2 (mul.rn.f32x2 / add.rn.f32x2 via fma identities)
// on sm_100a. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o packed_fp32 packed_fp32.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk(float a, float b) {
	unsigned long long r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
	return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
	unsigned long long r;
	asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
	unsigned long long r;
	asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}

template<int MODE>
__global__ void bench(float* out, int iters, float s) {
	float a[8], b[8];
	for (int i = 0; i < 8; i++) {
		a[i] = threadIdx.x * 0.001f + i;
		b[i] = 1.0f + i * 1e-3f;
	}
	if (MODE == 0) {  // scalar mul then add chains, 8 independent
		for (int it = 0; it < iters; it++) {
#pragma unroll
			for (int i = 0; i < 8; i++) a[i] = a[i] * s;
#pragma unroll
			for (int i = 0; i < 8; i++) a[i] = a[i] + b[i];
		}
	} else if (MODE == 1) {  // packed
		unsigned long long p[4], q[4], ss = pk(s, s);
		for (int i = 0; i < 4; i++) {
			p[i] = pk(a[2 * i], a[2 * i + 1]);
			q[i] = pk(b[2 * i], b[2 * i + 1]);
		}
		for (int it = 0; it < iters; it++) {
#pragma unroll
			for (int i = 0; i < 4; i++) p[i] = mul2(p[i], ss);
#pragma unroll
			for (int i = 0; i < 4; i++) p[i] = add2(p[i], q[i]);
		}
		for (int i = 0; i < 4; i++) {
			float lo, hi;
			asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
			a[2 * i] = lo;
			a[2 * i + 1] = hi;
		}
	} else {  // scalar interleaved with integer work (alu pipe)
		int z[8];
		for (int i = 0; i < 8; i++) z[i] = threadIdx.x + i;
		for (int it = 0; it < iters; it++) {
#pragma unroll
			for (int i = 0; i < 8; i++) {
				a[i] = a[i] * s;
				z[i] = z[i] + (z[i] >> 3);
			}
		}
		for (int i = 0; i < 8; i++) a[i] += z[i];
	}
	float r = 0;
	for (int i = 0; i < 8; i++) r += a[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main() {
	float* out;
	cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
	const int iters = 20000;
	for (int mode = 0; mode < 3; mode++) {
		cudaEvent_t a, b;
		cudaEventCreate(&a);
		cudaEventCreate(&b);
		for (int rep = 0; rep < 2; rep++) {
			cudaEventRecord(a);
			if (mode == 0) bench<0><<<148 * 8, 256>>>(out, iters, 1.0000001f);
			if (mode == 1) bench<1><<<148 * 8, 256>>>(out, iters, 1.0000001f);
			if (mode == 2) bench<2><<<148 * 8, 256>>>(out, iters, 1.0000001f);
			cudaEventRecord(b);
			cudaEventSynchronize(b);
		}
		float ms;
		cudaEventElapsedTime(&ms, a, b);
		// flops per thread: mode 0/1: 16 per iteration (8 mul + 8 add); mode 2: 8 mul + 16 int
		const double warps = 148.0 * 8 * 8, ops = mode == 2 ? 8 : 16;
		const double per_smsp_cycles = ms * 1e-3 * 1.965e9;  // assumes max clock
		const double warp_ops_per_smsp = warps / (148 * 4) * iters * ops;
		printf("mode %d: %.3f ms, %.2f cycles per warp-wide fp32 op per SMSP (%.1f TFLOP/s of mul/add)\n", mode, ms,
				per_smsp_cycles / warp_ops_per_smsp, 148.0 * 8 * 256 * iters * ops / (ms * 1e-3) / 1e12);
	}
	return 0;
}
