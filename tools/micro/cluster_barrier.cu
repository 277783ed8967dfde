// cluster_barrier.cu -- what does a barrier of one thread-block cluster cost on B200? (floor of the single-cluster 2D optimizers,
// csrc/slavcheva_persistent.cu / csrc/hier2d_persistent.cu). One cluster of B blocks x T threads runs N iterations of
//   mode 0: cluster.sync()
//   mode 1: one store into the next block's shared memory per thread, then cluster.sync()
//   mode 2: block-wide maximum (__syncthreads + shuffles), one store per block into every block's slot array, cluster.sync(),
//           every thread reads the slots (the termination test of the optimizers)
//   mode 3: __syncthreads() only (no cluster barrier)
// and the host prints the time per iteration. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_barrier
// tools/micro/cluster_barrier.cu; run: ./cluster_barrier
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>

namespace cg = cooperative_groups;

__global__ void __launch_bounds__(1024, 1) k_barriers(int mode, int iterations, float* out) {
	__shared__ float slots[16], warp_max[32], cell[1024];
	cg::cluster_group cluster = cg::this_cluster();
	const int rank = (int) cluster.block_rank(), blocks = (int) cluster.num_blocks();
	float value = (float) threadIdx.x;
	cell[threadIdx.x] = 0.0f;
	cluster.sync();
	for (int i = 0; i < iterations; i++) {
		if (mode == 0) cluster.sync();
		else if (mode == 1) {
			*cluster.map_shared_rank(&cell[threadIdx.x], (rank + 1) % blocks) = value;
			cluster.sync();
			value += cell[threadIdx.x];
		} else if (mode == 2) {
			float v = value;
			for (int offset = 16; offset > 0; offset >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, offset));
			if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = v;
			__syncthreads();
			if (threadIdx.x < 32) {
				v = threadIdx.x < (blockDim.x >> 5) ? warp_max[threadIdx.x] : 0.0f;
				for (int offset = 16; offset > 0; offset >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, offset));
			}
			if (threadIdx.x < blocks) *cluster.map_shared_rank(&slots[rank], threadIdx.x) = v;
			cluster.sync();
			float m = 0.0f;
			for (int k = 0; k < blocks; k++) m = fmaxf(m, slots[k]);
			value = m * 0.5f + 1.0f;
		} else {
			__syncthreads();
			value += 1.0f;
		}
	}
	if (out != nullptr) out[blockIdx.x * blockDim.x + threadIdx.x] = value;
}

int main() {
	float* out;
	cudaMalloc(&out, 16 * 1024 * sizeof(float));
	cudaFuncSetAttribute(k_barriers, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
	cudaEvent_t start, stop;
	cudaEventCreate(&start);
	cudaEventCreate(&stop);
	const int iterations = 2000;
	const int block_counts[] = { 1, 2, 4, 8, 16 }, thread_counts[] = { 128, 256, 1024 };
	for (int mode = 0; mode < 4; mode++)
		for (int blocks : block_counts)
			for (int threads : thread_counts) {
				cudaLaunchConfig_t config = {};
				config.gridDim = dim3(blocks);
				config.blockDim = dim3(threads);
				cudaLaunchAttribute attribute;
				attribute.id = cudaLaunchAttributeClusterDimension;
				attribute.val.clusterDim.x = blocks;
				attribute.val.clusterDim.y = attribute.val.clusterDim.z = 1;
				config.attrs = &attribute;
				config.numAttrs = 1;
				float ms[2] = { 0.0f, 0.0f };
				bool ok = true;
				for (int pass = 0; pass < 2 && ok; pass++) {  // 1 x and 2 x the iterations: the difference removes the launch
					for (int repeat = 0; repeat < 2; repeat++) {
						cudaEventRecord(start);
						ok = cudaLaunchKernelEx(&config, k_barriers, mode, iterations * (pass + 1), out) == cudaSuccess;
						cudaEventRecord(stop);
						ok = ok && cudaEventSynchronize(stop) == cudaSuccess;
						cudaEventElapsedTime(&ms[pass], start, stop);
					}
				}
				if (!ok) {
					printf("mode %d blocks %2d threads %4d: launch failed (%s)\n", mode, blocks, threads, cudaGetErrorString(cudaGetLastError()));
					continue;
				}
				printf("mode %d blocks %2d threads %4d: %.0f ns per iteration\n", mode, blocks, threads, 1e6 * (ms[1] - ms[0]) / iterations);
			}
	return 0;
}
