// slab_peer.cu -- halo exchange of the slab decomposition (BASELINE.json configs[4]) through peer memory:
// the ranks of one NVLink / NVSwitch domain write their boundary planes straight into the neighbours' halo planes.
//
// One process per GPU. Every rank owns ONE allocation made with cudaMalloc (exportable by cudaIpcGetMemHandle) that holds
// the two exchanged gradient fields of the current pyramid level at fixed offsets and a mailbox; every rank maps the
// allocations of all other ranks (cudaIpcOpenMemHandle enables peer access). Per exchange ONE kernel runs on each rank:
//   1. all blocks copy the rank's first / last `width` owned planes of the field into the low / high neighbour's halo planes
//      (128-bit stores over NVLink; a component's planes are contiguous), then fence at system scope;
//   2. the last block to finish signals: it stores the exchange's sequence number into the neighbours' mailboxes and, when the
//      exchange also carries the level-termination reduction (optimizer.tpp:166-171), the rank's max ||g||^2 bits tagged with
//      the sequence number into every rank's mailbox;
//   3. the same block waits until its own mailbox shows the neighbours' signals (and all ranks' maxima), writes the reduced
//      maximum into the iteration's slot and ends the kernel -- the next phase kernel in the stream then finds its halo
//      planes complete. No host round trip, no NCCL call, no staging copy: 1 launch per exchange.
// Ordering argument (why a neighbour's store can never hit planes that are still being read, and why the buffers need no
// second copy) is in slab.py (PeerExchange). A wait that lasts longer than ~4 s sets the mailbox's error flag and ends the
// kernel (the host raises), so a rank that died cannot hang the others' GPUs.
// Virtual ranks of one process (tests on a single GPU) use the same kernel with plain device pointers and one stream per
// rank.
#include "common.cuh"

#include <algorithm>
#include <cstring>

namespace lsf {
namespace {

constexpr int MAX_PEERS = LSF_SLAB_MAX_PEERS;
constexpr int MAX_RING = 4;

// mailbox of a rank, written by the other ranks (system-scope stores), read by the rank's waiting block
struct Mailbox {
	unsigned from_low;                             // sequence number of the last exchange whose planes the low neighbour delivered
	unsigned from_high;
	unsigned long long maxima[MAX_RING][MAX_PEERS];  // (sequence << 32) | bits of a rank's max ||g||^2
	unsigned arrivals;                             // local: blocks of the running exchange kernel that finished copying
	int error;                                     // local: 1 = a wait timed out
};
static_assert(sizeof(Mailbox) <= LSF_SLAB_MAILBOX_BYTES, "mailbox does not fit");

struct ExchangeArgs {
	const float* field;       // [3][planes][Y][Z] of this rank
	float* low_field;         // the low / high neighbour's field of the same level (nullptr at the volume border)
	float* high_field;
	long long plane;          // Y * Z
	long long planes, low_planes, high_planes;  // allocation planes: component stride / plane
	int own_begin, own_end;   // this rank's owned planes
	int low_dst, high_dst;    // first halo plane in the low neighbour's (its high halo) / high neighbour's (its low halo) allocation
	int width;                // planes per direction (0: reduction only)
	Mailbox* mailbox;         // own
	Mailbox* peer_mailbox[MAX_PEERS];
	int rank, world_size;
	unsigned sequence;
	unsigned* slot;           // max slot of the iteration to reduce over the ranks, or nullptr
};

__device__ __forceinline__ void store_system(unsigned* p, unsigned v) {
	asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void store_system(unsigned long long* p, unsigned long long v) {
	asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned load_system(const unsigned* p) {
	unsigned v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ unsigned long long load_system(const unsigned long long* p) {
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void copy_planes(const float* __restrict__ src, float* __restrict__ dst, long long count,
		long long first, long long stride) {
	if ((count & 3) == 0 && ((reinterpret_cast<unsigned long long>(src) | reinterpret_cast<unsigned long long>(dst)) & 15) == 0) {
		const float4* s = reinterpret_cast<const float4*>(src);
		float4* d = reinterpret_cast<float4*>(dst);
		for (long long i = first; i < (count >> 2); i += stride) d[i] = s[i];
	} else {
		for (long long i = first; i < count; i += stride) dst[i] = src[i];
	}
}

__global__ void __launch_bounds__(256) k_slab_exchange(const ExchangeArgs a) {
	const long long first = (long long) blockIdx.x * blockDim.x + threadIdx.x, stride = (long long) gridDim.x * blockDim.x;
	const long long count = a.width * a.plane;
	if (count > 0) {
		for (int c = 0; c < 3; c++) {
			if (a.low_field)
				copy_planes(a.field + (c * a.planes + a.own_begin) * a.plane, a.low_field + (c * a.low_planes + a.low_dst) * a.plane,
						count, first, stride);
			if (a.high_field)
				copy_planes(a.field + (c * a.planes + a.own_end - a.width) * a.plane,
						a.high_field + (c * a.high_planes + a.high_dst) * a.plane, count, first, stride);
		}
	}
	__threadfence_system();
	__shared__ bool last;
	__syncthreads();
	if (threadIdx.x == 0) last = atomicAdd(&a.mailbox->arrivals, 1u) == gridDim.x - 1;
	__syncthreads();
	if (!last) return;
	__threadfence_system();
	// ---- signal
	if (threadIdx.x == 0) {
		a.mailbox->arrivals = 0;
		if (count > 0) {
			if (a.low_field) store_system(&a.peer_mailbox[a.rank - 1]->from_high, a.sequence);
			if (a.high_field) store_system(&a.peer_mailbox[a.rank + 1]->from_low, a.sequence);
		}
	}
	if (a.slot && (int) threadIdx.x < a.world_size) {
		const unsigned long long tagged = ((unsigned long long) a.sequence << 32) | (unsigned long long) *a.slot;
		store_system(&a.peer_mailbox[threadIdx.x]->maxima[a.sequence % MAX_RING][a.rank], tagged);
	}
	// ---- wait (thread t < world_size: rank t's maximum; thread 32 / 33: the neighbours' planes)
	const long long start = clock64();
	// ~4 s at 2 GHz; once a wait has timed out the following exchanges do not wait at all (the run is lost, the host raises)
	const long long limit = a.mailbox->error ? 0 : 8000000000ll;
	unsigned mine = 0;
	bool timed_out = false;
	if (a.slot && (int) threadIdx.x < a.world_size) {
		const unsigned long long* entry = &a.mailbox->maxima[a.sequence % MAX_RING][threadIdx.x];
		unsigned long long v;
		while ((unsigned) ((v = load_system(entry)) >> 32) != a.sequence) {
			if (clock64() - start > limit) {
				timed_out = true;
				break;
			}
			__nanosleep(40);
		}
		mine = (unsigned) v;
	}
	if (count > 0 && ((threadIdx.x == 32 && a.low_field) || (threadIdx.x == 33 && a.high_field))) {
		const unsigned* flag = threadIdx.x == 32 ? &a.mailbox->from_low : &a.mailbox->from_high;
		// sequence numbers only grow; a neighbour may already be one exchange ahead (its next planes go to the other field)
		while ((int) (load_system(flag) - a.sequence) < 0) {
			if (clock64() - start > limit) {
				timed_out = true;
				break;
			}
			__nanosleep(40);
		}
	}
	if (timed_out) a.mailbox->error = 1;
	__shared__ unsigned reduced;
	if (threadIdx.x == 0) reduced = 0;
	__syncthreads();
	if (a.slot && (int) threadIdx.x < a.world_size) atomicMax(&reduced, mine);  // bits of non-negative floats order like integers
	__syncthreads();
	if (a.slot && threadIdx.x == 0) *a.slot = reduced;
	__threadfence_system();
}

}  // namespace
}  // namespace lsf

using namespace lsf;

extern "C" int lsf_peer_alloc(size_t bytes, void** pointer_out, unsigned char* handle_out) {
	LSF_REQUIRE(bytes > 0 && pointer_out, "bytes must be positive and pointer_out must not be NULL");
	void* p = nullptr;
	LSF_CUDA(cudaMalloc(&p, bytes));
	LSF_CUDA(cudaMemset(p, 0, bytes));
	if (handle_out) {
		cudaIpcMemHandle_t handle;
		static_assert(sizeof(handle) == LSF_PEER_HANDLE_BYTES, "handle size");
		const cudaError_t status = cudaIpcGetMemHandle(&handle, p);
		if (status != cudaSuccess) {
			cudaFree(p);
			LSF_CUDA(status);
		}
		memcpy(handle_out, &handle, sizeof(handle));
	}
	*pointer_out = p;
	return LSF_OK;
}

extern "C" int lsf_peer_open(const unsigned char* handle_in, void** pointer_out) {
	LSF_REQUIRE(handle_in && pointer_out, "handle and pointer_out must not be NULL");
	cudaIpcMemHandle_t handle;
	memcpy(&handle, handle_in, sizeof(handle));
	LSF_CUDA(cudaIpcOpenMemHandle(pointer_out, handle, cudaIpcMemLazyEnablePeerAccess));
	return LSF_OK;
}

extern "C" int lsf_peer_close(void* pointer) {
	if (pointer) LSF_CUDA(cudaIpcCloseMemHandle(pointer));
	return LSF_OK;
}

extern "C" int lsf_peer_free(void* pointer) {
	if (pointer) LSF_CUDA(cudaFree(pointer));
	return LSF_OK;
}

extern "C" int lsf_slab_exchange(const lsf_slab_peers* peers, const lsf_slab_level* level, size_t field_offset, int width,
		int low_planes, int low_destination_plane, int high_planes, int high_destination_plane, int reduce_iteration,
		unsigned sequence, void* stream_handle) {
	cudaStream_t stream = static_cast<cudaStream_t>(stream_handle);
	LSF_REQUIRE(peers && level, "peers and level must not be NULL");
	LSF_REQUIRE(peers->world_size >= 1 && peers->world_size <= MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world_size,
			"rank %d of %d ranks: at most %d ranks are supported", peers->rank, peers->world_size, MAX_PEERS);
	LSF_REQUIRE(width >= 0 && width <= level->own_end - level->own_begin, "width %d exceeds the owned planes", width);
	LSF_REQUIRE(sequence > 0, "sequence numbers start at 1");
	for (int r = 0; r < peers->world_size; r++) LSF_REQUIRE(peers->base[r], "allocation of rank %d is not mapped", r);
	const bool low = peers->rank > 0, high = peers->rank < peers->world_size - 1;
	if (width > 0) {
		LSF_REQUIRE(!low || (low_destination_plane >= 0 && low_destination_plane + width <= low_planes),
				"invalid halo planes of the low neighbour");
		LSF_REQUIRE(!high || (high_destination_plane >= 0 && high_destination_plane + width <= high_planes),
				"invalid halo planes of the high neighbour");
	}
	ExchangeArgs a;
	char* own = static_cast<char*>(peers->base[peers->rank]);
	a.field = reinterpret_cast<const float*>(own + field_offset);
	a.low_field = low && width > 0 ? reinterpret_cast<float*>(static_cast<char*>(peers->base[peers->rank - 1]) + field_offset) : nullptr;
	a.high_field = high && width > 0 ? reinterpret_cast<float*>(static_cast<char*>(peers->base[peers->rank + 1]) + field_offset) : nullptr;
	a.plane = (long long) level->Y * level->Z;
	a.planes = level->planes;
	a.low_planes = low_planes;
	a.high_planes = high_planes;
	a.own_begin = level->own_begin;
	a.own_end = level->own_end;
	a.low_dst = low_destination_plane;
	a.high_dst = high_destination_plane;
	a.width = width;
	for (int r = 0; r < MAX_PEERS; r++)
		a.peer_mailbox[r] = r < peers->world_size
				? reinterpret_cast<Mailbox*>(static_cast<char*>(peers->base[r]) + peers->mailbox_offset) : nullptr;
	a.mailbox = a.peer_mailbox[peers->rank];
	a.rank = peers->rank;
	a.world_size = peers->world_size;
	a.sequence = sequence;
	a.slot = reduce_iteration >= 0 ? level->max_sq_bits + reduce_iteration : nullptr;
	if (!a.slot && !a.low_field && !a.high_field) return LSF_OK;  // nothing to send, nothing to wait for
	const long long vectors = (long long) width * a.plane / 4;
	// enough blocks to fill the NVLink ports (a few hundred KB in flight), one block for the reduction alone
	const unsigned blocks = (unsigned) std::min<long long>(std::max<long long>(div_up(vectors, 256 * 8), 1), 4 * 148);
	k_slab_exchange<<<counted(blocks), 256, 0, stream>>>(a);
	LSF_CUDA(cudaGetLastError());
	return LSF_OK;
}

extern "C" int lsf_slab_exchange_error(const lsf_slab_peers* peers, int* error_out, void* stream_handle) {
	cudaStream_t stream = static_cast<cudaStream_t>(stream_handle);
	LSF_REQUIRE(peers && error_out, "peers and error_out must not be NULL");
	const Mailbox* mailbox = reinterpret_cast<const Mailbox*>(static_cast<char*>(peers->base[peers->rank]) + peers->mailbox_offset);
	LSF_CUDA(cudaMemcpyAsync(error_out, &mailbox->error, sizeof(int), cudaMemcpyDeviceToHost, stream));
	LSF_CUDA(cudaStreamSynchronize(stream));
	return LSF_OK;
}

extern "C" int lsf_hier_slab_iterations(const lsf_hier_params* params, const lsf_slab_level* level_in,
		const lsf_slab_peers* peers, const lsf_slab_link* link, int first_iteration, int iteration_count,
		unsigned first_sequence, unsigned* sequences_used, void* stream) {
	LSF_REQUIRE(params && level_in && peers && link, "params, level, peers and link must not be NULL");
	LSF_REQUIRE(first_iteration >= 0 && iteration_count >= 0, "invalid iteration range");
	const bool tikhonov = params->tikhonov_term_enabled && params->tikhonov_strength > 0.0f;
	const bool use_kernel = params->gradient_kernel_enabled && params->kernel_size > 0 && params->kernel != nullptr;
	const int radius = use_kernel ? params->kernel_size / 2 : 0;
	lsf_slab_level level = *level_in;
	size_t pre_offset = link->pre_offset, post_offset = link->post_offset;
	const char* own = static_cast<const char*>(peers->base[peers->rank]);
	LSF_REQUIRE(reinterpret_cast<const char*>(level.g_pre) == own + pre_offset
			&& reinterpret_cast<const char*>(level.g_post) == own + post_offset,
			"g_pre / g_post of the level are not at the link's offsets inside the rank's allocation");
	unsigned sequence = first_sequence;
	for (int it = first_iteration; it < first_iteration + iteration_count; it++) {
		LSF_TRY(lsf_hier_slab_iteration(params, &level, it, 1, stream));
		if (use_kernel) {
			LSF_TRY(lsf_slab_exchange(peers, &level, pre_offset, radius, link->low_planes, link->low_own_end, link->high_planes,
					link->high_own_begin - radius, -1, sequence++, stream));
			LSF_TRY(lsf_hier_slab_iteration(params, &level, it, 2, stream));
		} else {  // phase 1 wrote the iteration's gradient into g_pre: it is the next iteration's g_post
			std::swap(level.g_pre, level.g_post);
			std::swap(pre_offset, post_offset);
		}
		const int width = tikhonov ? 1 : 0;
		LSF_TRY(lsf_slab_exchange(peers, &level, post_offset, width, link->low_planes, link->low_own_end, link->high_planes,
				link->high_own_begin - width, it, sequence++, stream));
	}
	if (sequences_used) *sequences_used = sequence - first_sequence;
	return LSF_OK;
}
