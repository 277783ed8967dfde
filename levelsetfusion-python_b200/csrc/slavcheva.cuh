// slavcheva.cuh -- sm_100a kernels of the SobolevFusion / KillingFusion ("slavcheva") optimizers, 2D and 3D.
//
// One iteration (reference SobolevOptimizer2d::perform_optimization_iteration_and_return_max_warp,
// cpp/src/nonrigid_optimization/slavcheva/sobolev_optimizer2d.cpp:121-138; Python twin
// nonrigid_opt/slavcheva/slavcheva_optimizer2d.py:163-330) is
//   k_slav_gradient      live gradient + data term + Tikhonov/Killing term + level-set term + band-union mask + weights
//   k_slav_filter_axis   one pass of the separable Sobolev filter with the preserve-zeros rule (x D)
//   k_slav_resample      re-warp of the live field by the new warp (masks, truncation snap, warp zeroing) + max ||warp||
//   k_slav_decide        one thread: evaluates the termination test on the device (sticky), so the host polls once per
//                        chunk of iterations
// Device layout: scalar fields f[n0][n1][n2] (last axis contiguous), vector fields as D planes p[c][N].
// Component c displaces along array axis comp_axis(c): 2D {1,0} (u -> columns, v -> rows), 3D {0,1,2}.
//
// Three semantics (see include/lsf_b200.h): CPP = the C++ SobolevOptimizer2d (and its 3D / Killing / level-set
// generalisation), PY_DIRECT / PY_VECTORIZED = the two compute methods of the Python SlavchevaOptimizer2d. All
// arithmetic is float32 in the reference's operation order, no FMA (compiled with --fmad=false); the Python DIRECT
// re-warp interpolates in float64 like the reference's Python does.
#pragma once

#include "common.cuh"

namespace lsf {

struct SlavGeom {
	int nd;
	int n[3];
	int stride[3];
	int comp_axis[3];
	long long N;
};

inline SlavGeom make_slav_geom(int nd, const int* dims) {
	SlavGeom g;
	g.nd = nd;
	g.N = 1;
	for (int a = 0; a < 3; a++) g.n[a] = a < nd ? dims[a] : 1;
	for (int a = 2; a >= 0; a--) {
		if (a < nd) {
			g.stride[a] = (int) g.N;
			g.N *= g.n[a];
		} else g.stride[a] = 0;
	}
	if (nd == 2) {
		g.comp_axis[0] = 1;
		g.comp_axis[1] = 0;
		g.comp_axis[2] = 0;
	} else {
		g.comp_axis[0] = 0;
		g.comp_axis[1] = 1;
		g.comp_axis[2] = 2;
	}
	return g;
}

struct SlavParams {
	int semantics, data_term_method, smoothing_term_method, level_set;
	float rate, data_weight, smoothing_weight, lambda, level_set_weight;
	float lower, upper;
	int min_iterations;
};

struct SlavGradientArgs {
	SlavGeom g;
	SlavParams p;
	const float* live;
	const float* canonical;
	const float* warp;    // planes, previous iteration (after zeroing)
	const float* stale;   // planes: gradient field of the previous iteration (PY_DIRECT), else nullptr
	float* out;           // planes
	const int* status;
	int iteration;
};

struct SlavFilterArgs {
	SlavGeom g;
	const float* in;        // planes
	float* out;             // planes
	const float* original;  // planes: the field before the first pass (PY zero rule), else nullptr
	float k[LSF_MAX_KERNEL_SIZE];
	int size, radius;
	int axis;
	int zero_rule;          // 0 none, 1 C++ (pass-input vector exactly zero), 2 Python (|original component| < 1e-6)
	const int* status;
	int iteration;
};

struct SlavResampleArgs {
	SlavGeom g;
	SlavParams p;
	const float* live;
	const float* canonical;
	const float* update;     // planes: filtered field (CPP: the new warp; PY: the gradient)
	float* gradient_field;   // planes, zeroed where the value snaps (PY_DIRECT: aliases update), else nullptr
	float* warp;             // planes out
	float* new_live;
	int band_union_only, known_values_only, substitute_original, modify_warp;
	float threshold;
	unsigned* max_sq_bits;   // slot of this iteration (nullptr: no reduction)
	const int* status;
	int iteration;
};

// One iteration of the generic (dense, dimension-generic) kernel sequence as data: what k_slav_gradient, the
// k_slav_filter_axis passes and k_slav_resample would be launched with. Small fields run a whole polling chunk of
// iterations in ONE cooperative launch that reads these records (slavcheva_persistent.cu).
struct SlavIterationCommand {
	SlavGradientArgs gradient;
	SlavFilterArgs pass[3];
	int passes;
	SlavResampleArgs resample;
};
// largest field (voxels) the single-launch path takes: all blocks must be resident at once
long long slav_persistent_capacity();
// iterations [first_iteration, first_iteration + count) from `commands_dev`; the termination test (k_slav_decide) runs inside
// `buffers` (optional): every buffer the command records name -- fields of up to ~16 K voxels are then kept in the shared
// memory of one thread-block cluster for the whole launch (k_slav_strips)
struct SlavOptimizerBuffers {
	float* vector_fields[4];  // warp and the three update fields (planes)
	float* scalar_fields[3];  // the two live buffers and the canonical field (last: never written)
	int H, W, radius;         // field shape, filter radius (0: no filter)
};
int launch_slav_persistent2d(const SlavIterationCommand* commands_dev, int count, const SlavParams& p, long long N,
		unsigned* max_sq_bits, int* status, int first_iteration, int max_iterations, cudaStream_t stream,
		const SlavOptimizerBuffers* buffers = nullptr);

// Band-union warp statistics and TSDF difference statistics of device-resident fields (defined in slavcheva.cu; also
// used by the hierarchical optimizers for their per-level convergence reports). `field` element (voxel i,
// component c) = field[c * component_stride + i * voxel_stride]. Either output may be NULL.
int statistics_on_device(int nd, const int* dims, const float* field, long long component_stride, int voxel_stride,
		const float* canonical, const float* live, float min_threshold, float max_threshold,
		lsf_warp_delta_statistics_t* warp_out, lsf_tsdf_difference_statistics_t* diff_out, Arena& arena,
		cudaStream_t stream);

#ifdef __CUDACC__

__device__ __forceinline__ bool slav_truncated(float v) {
	return fabsf(v) == 1.0f;  // reference boolean_operations.hpp:37-39
}

template<int D>
__device__ __forceinline__ void slav_coords(const SlavGeom& g, int idx, int (&pos)[3]) {
	pos[0] = pos[1] = pos[2] = 0;
#pragma unroll
	for (int a = 0; a < D; a++) {
		pos[a] = idx / g.stride[a];
		idx -= pos[a] * g.stride[a];
	}
}

template<int D>
__device__ __forceinline__ bool slav_inside(const SlavGeom& g, const int (&q)[3]) {
	bool ok = true;
#pragma unroll
	for (int a = 0; a < D; a++) ok = ok && q[a] >= 0 && q[a] < g.n[a];
	return ok;
}

template<int D>
__device__ __forceinline__ int slav_index(const SlavGeom& g, const int (&q)[3]) {
	int idx = 0;
#pragma unroll
	for (int a = 0; a < D; a++) idx += q[a] * g.stride[a];
	return idx;
}

// IN = true: the caller knows that the voxel is an interior one (every stencil tap is inside the field), so the bounds
// tests drop out; 3D fields have component c along array axis c, so the axis look-ups fold away as well.
template<int D>
__device__ __forceinline__ int slav_axis(const SlavGeom& g, int component) {
	return D == 3 ? component : g.comp_axis[component];
}

// SM = true: the field pointers of the argument block address a tile staged in shared memory (its own strides in g)
template<bool SM>
__device__ __forceinline__ float slav_ld(const float* p) {
	if (SM) {
		__builtin_assume(__isShared(p));
		return *p;
	}
	return __ldg(p);
}

template<int D, bool IN = false, bool SM = false>
__device__ __forceinline__ float slav_live_or_one(const SlavGeom& g, const float* __restrict__ live, const int (&q)[3]) {
	return (IN || slav_inside<D>(g, q)) ? slav_ld<SM>(live + slav_index<D>(g, q)) : 1.0f;
}

// ---------------------------------------------------------------------------------------------- gradient terms
// data term: reference data_term.cpp:63-84 (C++), data_term.py:169-227 (Python basic / thresholded FDM)
template<int D, bool IN = false, bool SM = false>
__device__ __forceinline__ void slav_data_term_given(const SlavGradientArgs& a, int idx, const int (&pos)[3],
		float canonical_value, float (&out)[3]) {
	const SlavGeom& g = a.g;
	const float centre = slav_ld<SM>(a.live + idx);
	float grad[3] = { 0.f, 0.f, 0.f };
#pragma unroll
	for (int c = 0; c < D; c++) {
		const int ax = slav_axis<D>(g, c), i = pos[ax], n = g.n[ax], s = g.stride[ax];
		if (IN) {
			grad[c] = 0.5f * (slav_ld<SM>(a.live + idx + s) - slav_ld<SM>(a.live + idx - s));
		} else if (n >= 2) {
			if (i == 0) grad[c] = slav_ld<SM>(a.live + idx + s) - centre;
			else if (i == n - 1) grad[c] = centre - slav_ld<SM>(a.live + idx - s);
			else grad[c] = 0.5f * (slav_ld<SM>(a.live + idx + s) - slav_ld<SM>(a.live + idx - s));
		}
		if (a.p.data_term_method == LSF_DATA_TERM_THRESHOLDED_FDM && fabsf(grad[c]) > 0.5f) {
			const float minus = (IN || i > 0) ? slav_ld<SM>(a.live + idx - s) : 1.0f;
			const float plus = (IN || i < n - 1) ? slav_ld<SM>(a.live + idx + s) : 1.0f;
			const float forward = plus - centre, backward = centre - minus;
			float value = fabsf(forward) < fabsf(backward) ? forward : backward;
			if (fabsf(value) > 0.5f) value = 0.0f;
			grad[c] = value;
		}
	}
	const float diff = centre - canonical_value;
	if (a.p.semantics == LSF_SEMANTICS_CPP) {
		const float scaled = 10.0f * diff;
#pragma unroll
		for (int c = 0; c < D; c++) out[c] = scaled * grad[c];
	} else {
#pragma unroll
		for (int c = 0; c < D; c++) out[c] = (diff * grad[c]) * 10.0f;
	}
}

template<int D, bool IN = false>
__device__ __forceinline__ void slav_data_term(const SlavGradientArgs& a, int idx, const int (&pos)[3], float (&out)[3]) {
	slav_data_term_given<D, IN, false>(a, idx, pos, __ldg(a.canonical + idx), out);
}

template<int D, bool IN = false, bool SM = false>
__device__ __forceinline__ float slav_warp_or_centre(const SlavGradientArgs& a, const int (&q)[3], int c, int centre_idx) {
	const int at = (IN || slav_inside<D>(a.g, q)) ? slav_index<D>(a.g, q) : centre_idx;
	return slav_ld<SM>(a.warp + c * a.g.N + at);
}

// C++ Tikhonov term: reference smoothing_term.cpp:43-108 (array axis 0 assigned, further axes added)
template<int D, bool IN = false, bool SM = false>
__device__ __forceinline__ void slav_tikhonov_cpp(const SlavGradientArgs& a, int idx, const int (&pos)[3], float (&out)[3]) {
#pragma unroll
	for (int c = 0; c < D; c++) {
		const float* w = a.warp + c * a.g.N + idx;
		const float centre = slav_ld<SM>(w);
		float total = 0.0f;
#pragma unroll
		for (int ax = 0; ax < D; ax++) {
			const int i = pos[ax], n = a.g.n[ax], s = a.g.stride[ax];
			float term;
			if (!IN && n < 2) term = 0.0f;
			else if (!IN && i == 0) term = -slav_ld<SM>(w + s) + centre;
			else if (!IN && i == n - 1) term = -slav_ld<SM>(w - s) + centre;
			else term = (-slav_ld<SM>(w + s) + 2.0f * centre) - slav_ld<SM>(w - s);
			if (ax == 0) total = term;
			else total += term;
		}
		out[c] = total;
	}
}

// Python Tikhonov term: reference smoothing_term.py:103-139 (copy_if_zero=False)
template<int D>
__device__ __forceinline__ void slav_tikhonov_py(const SlavGradientArgs& a, int idx, const int (&pos)[3], float (&out)[3]) {
#pragma unroll
	for (int c = 0; c < D; c++) {
		int q[3] = { pos[0], pos[1], pos[2] };
		float acc = 0.0f;
#pragma unroll
		for (int k = 0; k < D; k++) {
			const int ax = a.g.comp_axis[k];
			q[ax] = pos[ax] + 1;
			const float v = slav_warp_or_centre<D>(a, q, c, idx);
			q[ax] = pos[ax];
			acc = k == 0 ? v : acc + v;
		}
		acc = acc - (2.0f * D) * __ldg(a.warp + c * a.g.N + idx);
#pragma unroll
		for (int k = 0; k < D; k++) {
			const int ax = a.g.comp_axis[k];
			q[ax] = pos[ax] - 1;
			acc = acc + slav_warp_or_centre<D>(a, q, c, idx);
			q[ax] = pos[ax];
		}
		out[c] = -acc;
	}
}

// Killing term: reference smoothing_term.py:50-100 (copy_if_zero=False), quirks kept; 3D form = the
// same expression pattern
template<int D, bool IN = false, bool SM = false>
__device__ __forceinline__ void slav_killing(const SlavGradientArgs& a, int idx, const int (&pos)[3], float (&out)[3]) {
	const float lambda = a.p.lambda;
	const float c0 = (float) (-2.0 * (1.0 + (double) lambda));
	const int ax0 = slav_axis<D>(a.g, 0);
#pragma unroll
	for (int ca = 0; ca < D; ca++) {
		int q[3] = { pos[0], pos[1], pos[2] };
		const float w = slav_ld<SM>(a.warp + ca * a.g.N + idx);
		q[ax0] = pos[ax0] + 1;
		const float xp = slav_warp_or_centre<D, IN, SM>(a, q, ca, idx);
		q[ax0] = pos[ax0] - 1;
		const float xm = slav_warp_or_centre<D, IN, SM>(a, q, ca, idx);
		q[ax0] = pos[ax0];
		float acc = c0 * ((xp - 2.0f * w) + xm);
#pragma unroll
		for (int k = 1; k < D; k++) {
			const int ax = slav_axis<D>(a.g, k);
			q[ax] = pos[ax] + 1;
			const float yp = slav_warp_or_centre<D, IN, SM>(a, q, ca, idx);
			q[ax] = pos[ax];
			acc = acc + ((yp - 2.0f * w) + yp);
		}
#pragma unroll
		for (int cb = 0; cb < D; cb++) {
			if (cb == ca) continue;
			const int first = slav_axis<D>(a.g, ca < cb ? ca : cb), second = slav_axis<D>(a.g, ca < cb ? cb : ca);
			float v[4];
			int k = 0;
#pragma unroll
			for (int s1 = 1; s1 >= -1; s1 -= 2)
#pragma unroll
				for (int s2 = 1; s2 >= -1; s2 -= 2) {
					q[first] = pos[first] + s1;
					q[second] = pos[second] + s2;
					v[k++] = slav_warp_or_centre<D, IN, SM>(a, q, cb, idx);
				}
			q[first] = pos[first];
			q[second] = pos[second];
			const float mixed = (((v[0] - v[1]) - v[2]) + v[3]) / 4.0f;
			acc = acc + lambda * mixed;
		}
		out[ca] = acc;
	}
}

// level-set term: reference level_set_term.py:28-64 (out-of-bounds -> 1, quirks kept)
template<int D, bool IN = false, bool SM = false>
__device__ __forceinline__ void slav_level_set(const SlavGradientArgs& a, int idx, const int (&pos)[3], float (&out)[3]) {
	const SlavGeom& g = a.g;
	const float centre = slav_ld<SM>(a.live + idx);
	float grad[3] = { 0.f, 0.f, 0.f }, hessian[3][3];
	int q[3] = { pos[0], pos[1], pos[2] };
#pragma unroll
	for (int c = 0; c < D; c++) {
		const int ax = slav_axis<D>(g, c);
		q[ax] = pos[ax] + 1;
		const float plus = slav_live_or_one<D, IN, SM>(g, a.live, q);
		q[ax] = pos[ax] - 1;
		const float minus = slav_live_or_one<D, IN, SM>(g, a.live, q);
		q[ax] = pos[ax];
		grad[c] = (0.5f * (plus - minus)) * 10.0f;
		hessian[c][c] = ((plus - 2.0f * centre) + plus) * 10.0f;
	}
#pragma unroll
	for (int c1 = 0; c1 < D; c1++)
#pragma unroll
		for (int c2 = c1 + 1; c2 < D; c2++) {
			const int a1 = slav_axis<D>(g, c1), a2 = slav_axis<D>(g, c2);
			float v[4];
			int k = 0;
#pragma unroll
			for (int s2 = 1; s2 >= -1; s2 -= 2)
#pragma unroll
				for (int s1 = 1; s1 >= -1; s1 -= 2) {
					q[a1] = pos[a1] + s1;
					q[a2] = pos[a2] + s2;
					v[k++] = slav_live_or_one<D, IN, SM>(g, a.live, q);
				}
			q[a1] = pos[a1];
			q[a2] = pos[a2];
			const float mixed = (0.25f * (((v[0] - v[1]) - v[2]) + v[3])) * 10.0f;
			hessian[c1][c2] = mixed;
			hessian[c2][c1] = mixed;
		}
	float sq = 0.0f;
#pragma unroll
	for (int c = 0; c < D; c++) sq += grad[c] * grad[c];
	const float length = sqrtf(sq);
	const float factor = (1.0f - length) / (length + 1e-5f);
#pragma unroll
	for (int ca = 0; ca < D; ca++) {
		float acc = hessian[ca][0] * grad[0];
#pragma unroll
		for (int cb = 1; cb < D; cb++) acc = acc + hessian[ca][cb] * grad[cb];
		out[ca] = factor * acc;
	}
}

// the gradient terms at voxel idx (body of k_slav_gradient; also called by the single-launch optimizer of small fields,
// slavcheva_persistent.cu)
// known_pos (optional): the voxel's coordinates, if the caller has them (saves the integer divisions of slav_coords)
template<int D>
__device__ __forceinline__ void slav_gradient_at(const SlavGradientArgs& a, int idx, const int* known_pos = nullptr) {
	int pos[3];
	if (known_pos != nullptr) {
#pragma unroll
		for (int ax = 0; ax < 3; ax++) pos[ax] = known_pos[ax];
	} else
		slav_coords<D>(a.g, idx, pos);
	const float live_value = __ldg(a.live + idx);
	const bool outside = slav_truncated(live_value) && slav_truncated(__ldg(a.canonical + idx));
	const SlavParams& p = a.p;
	float result[3] = { 0.f, 0.f, 0.f };
	float data[3], smooth[3], ls[3];
	if (p.semantics == LSF_SEMANTICS_CPP) {
		if (outside) {
#pragma unroll
			for (int c = 0; c < D; c++) result[c] = (0.0f + 0.0f * p.smoothing_weight) * -p.rate;
		} else {
			bool interior = D == 3;  // 3D: interior voxels take the paths without bounds tests (same arithmetic)
#pragma unroll
			for (int ax = 0; ax < D; ax++) interior = interior && pos[ax] >= 1 && pos[ax] < a.g.n[ax] - 1;
			const bool ls_here = p.level_set && !slav_truncated(live_value);
			if (interior) {
				slav_data_term<D, true>(a, idx, pos, data);
				if (p.smoothing_term_method == LSF_SMOOTHING_KILLING) slav_killing<D, true>(a, idx, pos, smooth);
				else slav_tikhonov_cpp<D, true>(a, idx, pos, smooth);
				if (ls_here) slav_level_set<D, true>(a, idx, pos, ls);
			} else {
				slav_data_term<D>(a, idx, pos, data);
				if (p.smoothing_term_method == LSF_SMOOTHING_KILLING) slav_killing<D>(a, idx, pos, smooth);
				else slav_tikhonov_cpp<D>(a, idx, pos, smooth);
				if (ls_here) slav_level_set<D>(a, idx, pos, ls);
			}
#pragma unroll
			for (int c = 0; c < D; c++) {
				float total = data[c] * p.data_weight;
				if (ls_here) total = total + ls[c] * p.level_set_weight;
				total = total + smooth[c] * p.smoothing_weight;
				result[c] = total * -p.rate;  // reference sobolev_optimizer2d.cpp:131-132
			}
		}
	} else if (p.semantics == LSF_SEMANTICS_PY_DIRECT) {
		if (outside) {
			// reference slavcheva_optimizer2d.py:261-262: skipped voxels keep last iteration's gradient-field entry
#pragma unroll
			for (int c = 0; c < D; c++) result[c] = __ldg(a.stale + c * a.g.N + idx);
		} else {
			slav_data_term<D>(a, idx, pos, data);
			float total[3];
#pragma unroll
			for (int c = 0; c < D; c++) total[c] = 0.0f + p.data_weight * data[c];
			if (p.level_set && !slav_truncated(live_value)) {
				slav_level_set<D>(a, idx, pos, ls);
#pragma unroll
				for (int c = 0; c < D; c++) total[c] = total[c] + p.level_set_weight * ls[c];
			}
			if (p.smoothing_term_method == LSF_SMOOTHING_KILLING) slav_killing<D>(a, idx, pos, smooth);
			else slav_tikhonov_py<D>(a, idx, pos, smooth);
#pragma unroll
			for (int c = 0; c < D; c++) result[c] = total[c] + p.smoothing_weight * smooth[c];
		}
	} else {  // PY_VECTORIZED, reference slavcheva_optimizer2d.py:163-190
		if (!outside) {
			slav_data_term<D>(a, idx, pos, data);
			slav_tikhonov_cpp<D>(a, idx, pos, smooth);
#pragma unroll
			for (int c = 0; c < D; c++) result[c] = p.data_weight * data[c] + p.smoothing_weight * smooth[c];
		}
	}
#pragma unroll
	for (int c = 0; c < D; c++) a.out[c * a.g.N + idx] = result[c];
}

template<int D>
static __global__ void __launch_bounds__(256) k_slav_gradient(SlavGradientArgs a) {
	if (a.status[a.iteration]) return;
	const long long linear = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (linear >= a.g.N) return;
	slav_gradient_at<D>(a, (int) linear);
}

// The C++-semantics branch of k_slav_gradient<3> with four consecutive z voxels per thread (128-bit loads of the two
// fields and 128-bit stores of the result; the four voxels' band tests and stencils are independent instruction
// streams). Requires n[2] % 4 == 0 and 16-byte aligned fields. Same per-voxel functions, same arithmetic.
static __global__ void __launch_bounds__(256) k_slav_gradient_cpp3_v4(SlavGradientArgs a) {
	if (a.status[a.iteration]) return;
	const long long group = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (group * 4 >= a.g.N) return;
	const int base = (int) (group * 4);
	int pos[3];
	slav_coords<3>(a.g, base, pos);
	const SlavParams& p = a.p;
	const float4 live4 = __ldg(reinterpret_cast<const float4*>(a.live + base));
	const float4 canonical4 = __ldg(reinterpret_cast<const float4*>(a.canonical + base));
	const float live_v[4] = { live4.x, live4.y, live4.z, live4.w };
	const float canonical_v[4] = { canonical4.x, canonical4.y, canonical4.z, canonical4.w };
	float out[3][4];
	const bool killing = p.smoothing_term_method == LSF_SMOOTHING_KILLING;
	const bool xy_interior = pos[0] >= 1 && pos[0] < a.g.n[0] - 1 && pos[1] >= 1 && pos[1] < a.g.n[1] - 1;
#pragma unroll
	for (int v = 0; v < 4; v++) {
		const int idx = base + v;
		const int q[3] = { pos[0], pos[1], pos[2] + v };
		const float live_value = live_v[v];
		float result[3], data[3], smooth[3], ls[3];
		if (slav_truncated(live_value) && slav_truncated(canonical_v[v])) {
#pragma unroll
			for (int c = 0; c < 3; c++) result[c] = (0.0f + 0.0f * p.smoothing_weight) * -p.rate;
		} else {
			const bool ls_here = p.level_set && !slav_truncated(live_value);
			if (xy_interior && q[2] >= 1 && q[2] < a.g.n[2] - 1) {
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
				result[c] = total * -p.rate;  // reference sobolev_optimizer2d.cpp:131-132
			}
		}
#pragma unroll
		for (int c = 0; c < 3; c++) out[c][v] = result[c];
	}
#pragma unroll
	for (int c = 0; c < 3; c++)
		*reinterpret_cast<float4*>(a.out + c * a.g.N + base) = make_float4(out[c][0], out[c][1], out[c][2], out[c][3]);
}

// The same branch once more, with the narrow band compacted per block. Only the voxels inside the band union (17 % of
// a 256^3 sphere/plane pair) evaluate the stencils, and in k_slav_gradient_cpp3_v4 they sit in a few lanes of many
// warps: the warps run the long term code at a fraction of their width. Here a block of 256 threads owns 1024
// consecutive voxels; every thread classifies its four voxels, stores the constant result of the ones outside the band
// and appends the others to a list in shared memory; after a barrier the block works through the list with all its
// threads, one band voxel per thread and pass. Same per-voxel functions, same arithmetic, same results.
static __global__ void __launch_bounds__(256) k_slav_gradient_cpp3_band(SlavGradientArgs a) {
	if (a.status[a.iteration]) return;
	__shared__ unsigned short band_list[1024];
	__shared__ int band_count;
	const SlavParams& p = a.p;
	const long long block_base = (long long) blockIdx.x * 1024;
	if (threadIdx.x == 0) band_count = 0;
	__syncthreads();
	const long long first = block_base + threadIdx.x * 4;
	if (first < a.g.N) {
		const int base = (int) first;
		const float4 live4 = __ldg(reinterpret_cast<const float4*>(a.live + base));
		const float4 canonical4 = __ldg(reinterpret_cast<const float4*>(a.canonical + base));
		const float live_v[4] = { live4.x, live4.y, live4.z, live4.w };
		const float canonical_v[4] = { canonical4.x, canonical4.y, canonical4.z, canonical4.w };
		const float outside_value = (0.0f + 0.0f * p.smoothing_weight) * -p.rate;
		unsigned in_band = 0;
#pragma unroll
		for (int v = 0; v < 4; v++)
			if (!(slav_truncated(live_v[v]) && slav_truncated(canonical_v[v]))) in_band |= 1u << v;
		if (in_band != 0) {
			const int at = atomicAdd(&band_count, __popc(in_band));
			int k = 0;
#pragma unroll
			for (int v = 0; v < 4; v++)
				if (in_band & (1u << v)) band_list[at + k++] = (unsigned short) (threadIdx.x * 4 + v);
		}
		// the band voxels' slots are overwritten after the barrier
		const float4 constant = make_float4(outside_value, outside_value, outside_value, outside_value);
#pragma unroll
		for (int c = 0; c < 3; c++) *reinterpret_cast<float4*>(a.out + c * a.g.N + base) = constant;
	}
	__syncthreads();
	const int count = band_count;
	const bool killing = p.smoothing_term_method == LSF_SMOOTHING_KILLING;
	for (int j = threadIdx.x; j < count; j += 256) {
		const int idx = (int) block_base + band_list[j];
		int q[3];
		slav_coords<3>(a.g, idx, q);
		const float live_value = __ldg(a.live + idx);
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
}

// ---------------------------------------------------------------------------------------------- energy log (2D, Python)
// The {data, smoothing, level set} energy aggregates the reference's Python optimizer appends to its OptimizationLog
// every iteration (slavcheva_optimizer2d.py:370-374), for the state the iteration starts from: DIRECT :236-300 with the
// local energies of data_term.py:185,225, level_set_term.py:63, smoothing_term.py:97-98,137-138; VECTORIZED :163-175 with
// data_term.py:352-358 and smoothing_term.py:162-177 (np.gradient of the warp components). Evaluated in double from the
// float32 fields (like oracle slavcheva_energies); out[0..2] += this block's sums (double atomics: the order of the block
// sums is not fixed, the result is compared at 1e-9). C++ semantics keep no log: nothing is added.
static __global__ void __launch_bounds__(256) k_slav_energies2d(SlavGradientArgs a, double* __restrict__ out) {
	if (a.status[a.iteration]) return;
	const SlavGeom& g = a.g;
	const SlavParams& p = a.p;
	const int N = (int) g.N, H = g.n[0], W = g.n[1];
	double sums[3] = { 0.0, 0.0, 0.0 };
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx < N && p.semantics != LSF_SEMANTICS_CPP) {
		const float live_value = a.live[idx], canonical_value = a.canonical[idx];
		if (!(slav_truncated(live_value) && slav_truncated(canonical_value))) {
			const int row = idx / W, col = idx - row * W;
			const double diff = (double) live_value - (double) canonical_value;
			sums[0] = 0.5 * diff * diff;
			const float* u = a.warp;       // component 0 displaces along the columns ("x"), component 1 along the rows ("y")
			const float* v = a.warp + N;
			if (p.semantics == LSF_SEMANTICS_PY_DIRECT) {
				if (p.level_set && !slav_truncated(live_value)) {
					const double xp = col + 1 < W ? a.live[idx + 1] : 1.0f, xm = col >= 1 ? a.live[idx - 1] : 1.0f;
					const double yp = row + 1 < H ? a.live[idx + W] : 1.0f, ym = row >= 1 ? a.live[idx - W] : 1.0f;
					const double gx = 0.5 * (xp - xm) * 10.0, gy = 0.5 * (yp - ym) * 10.0;
					const double length = sqrt(gx * gx + gy * gy);
					sums[2] = 0.5 * (length - 1.0) * (length - 1.0);
				}
				// neighbours outside the field = the centre value (utils/sampling.py:84-88)
				const int xp = col + 1 < W ? idx + 1 : idx, xm = col >= 1 ? idx - 1 : idx;
				const int yp = row + 1 < H ? idx + W : idx, ym = row >= 1 ? idx - W : idx;
				const double ux = 0.5 * ((double) u[xp] - (double) u[xm]), vx = 0.5 * ((double) v[xp] - (double) v[xm]);
				const double uy = 0.5 * ((double) u[yp] - (double) u[ym]), vy = 0.5 * ((double) v[yp] - (double) v[ym]);
				if (p.smoothing_term_method == LSF_SMOOTHING_KILLING) {
					const double jj = ux * ux + vx * vx + uy * uy + vy * vy;
					const double jtj = ux * ux + uy * vx + vx * uy + vy * vy;
					sums[1] = jj + (double) p.lambda * jtj;
				} else {
					sums[1] = 0.5 * ((ux * ux + vx * vx) + (uy * uy + vy * vy));
				}
			} else {
				double aggregate = 0.0;
#pragma unroll
				for (int c = 0; c < 2; c++) {
					const float* w = c == 0 ? u : v;
					double d;
					if (row == 0) d = (double) w[idx + W] - (double) w[idx];
					else if (row == H - 1) d = (double) w[idx] - (double) w[idx - W];
					else d = 0.5 * ((double) w[idx + W] - (double) w[idx - W]);
					aggregate += d * d;
					if (col == 0) d = (double) w[idx + 1] - (double) w[idx];
					else if (col == W - 1) d = (double) w[idx] - (double) w[idx - 1];
					else d = 0.5 * ((double) w[idx + 1] - (double) w[idx - 1]);
					aggregate += d * d;
				}
				sums[1] = 0.5 * aggregate;
			}
		}
	}
	__shared__ double block_sums[8][3];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int k = 0; k < 3; k++) {
		double value = sums[k];
#pragma unroll
		for (int offset = 16; offset > 0; offset >>= 1) value += __shfl_xor_sync(0xffffffffu, value, offset);
		if (lane == 0) block_sums[warp][k] = value;
	}
	__syncthreads();
	if (threadIdx.x < 3) {
		double total = 0.0;
		for (int w = 0; w < 8; w++) total += block_sums[w][threadIdx.x];
		const double weight = threadIdx.x == 0 ? p.data_weight : (threadIdx.x == 1 ? p.smoothing_weight : p.level_set_weight);
		if (total != 0.0) atomicAdd(out + threadIdx.x, total * weight);
	}
}

// ---------------------------------------------------------------------------------------------- Sobolev filter pass
// reference convolve_with_kernel_preserve_zeros, cpp/src/math/convolution.cpp:23-67,69-145 (C++ rule) and
// math_utils/convolution.py:114-132 (Python rule)
// R > 0: the caller knows the filter radius (a.radius == R, a.size == 2R + 1): the tap loop unrolls and voxels at least R
// from both ends of their line skip the bounds tests (same taps, same order)
template<int D, int R = 0>
__device__ __forceinline__ void slav_filter_axis_at(const SlavFilterArgs& a, int idx, const int* known_pos = nullptr) {
	int pos[3];
	if (known_pos != nullptr) {
#pragma unroll
		for (int ax = 0; ax < 3; ax++) pos[ax] = known_pos[ax];
	} else
		slav_coords<D>(a.g, idx, pos);
	const int i = a.axis == 0 ? pos[0] : (a.axis == 1 ? pos[1] : pos[2]), n = a.g.n[a.axis], s = a.g.stride[a.axis];
	if (a.zero_rule == 1) {
		bool all_zero = true;
#pragma unroll
		for (int c = 0; c < D; c++) all_zero = all_zero && __ldg(a.in + c * a.g.N + idx) == 0.0f;
		if (all_zero) {
#pragma unroll
			for (int c = 0; c < D; c++) a.out[c * a.g.N + idx] = 0.0f;
			return;
		}
	}
#pragma unroll
	for (int c = 0; c < D; c++) {
		const float* line = a.in + c * a.g.N + idx;
		float acc = 0.0f;
		if (R > 0 && i >= R && i + R < n) {
#pragma unroll
			for (int j = 0; j < 2 * R + 1; j++) acc += __ldg(line + (j - R) * s) * a.k[j];
		} else if (R > 0) {
#pragma unroll
			for (int j = 0; j < 2 * R + 1; j++) {
				const int src = i - R + j;
				const float value = (src >= 0 && src < n) ? __ldg(line + (j - R) * s) : 0.0f;
				acc += value * a.k[j];
			}
		} else {
			for (int j = 0; j < a.size; j++) {
				const int src = i - a.radius + j;
				const float value = (src >= 0 && src < n) ? __ldg(line + (j - a.radius) * s) : 0.0f;
				acc += value * a.k[j];
			}
		}
		if (a.zero_rule == 2 && fabsf(__ldg(a.original + c * a.g.N + idx)) < 1e-6f) acc = 0.0f;
		a.out[c * a.g.N + idx] = acc;
	}
}

template<int D>
static __global__ void __launch_bounds__(256) k_slav_filter_axis(SlavFilterArgs a) {
	if (a.status[a.iteration]) return;
	const long long linear = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (linear >= a.g.N) return;
	slav_filter_axis_at<D>(a, (int) linear);
}

// ---------------------------------------------------------------------------------------------- re-warp of the live field
// reference warp_2d_advanced, cpp/src/nonrigid_optimization/field_warping.cpp:64-136 (float32), and its Python twin
// nonrigid_opt/field_warping.py:112-151 + utils/sampling.py:160-215 (float64 interpolation), followed by the
// maximum warp length (statistics.tpp:57-100; slavcheva_optimizer2d.py:309-318 measures BEFORE the re-warp).
// one voxel of the re-warp: `update` = its filtered update vector, live_value / canonical_value = the fields at the voxel;
// returns the new live value, the voxel's new warp vector and its contribution to the maximum warp length
// a box of the live field staged in shared memory: voxel q sits at data[((q0 - lo0) * ext1 + (q1 - lo1)) * ext2 + (q2 - lo2)]
struct SlavLiveTile {
	const float* data;
	int lo[3], ext[3];
};

// a tap of the re-warp's gather (live field at voxel q, row q[0], linear index): a plain load; the single-cluster optimizer of
// small 2D fields (slavcheva_persistent.cu) keeps the field in distributed shared memory and defines its own look-up
#ifndef SLAV_LIVE_TAP
#define SLAV_LIVE_TAP(live, row, index) __ldg((live) + (index))
#endif

// TILE: the eight taps are read from `tile` when the whole 2 x 2 x 2 cell lies inside it and inside the field (3D,
// float32 semantics only); other voxels take the global path
template<int D, bool TILE = false>
__device__ __forceinline__ void slav_resample_voxel(const SlavResampleArgs& a, int idx, const float (&update)[3],
		float live_value, float canonical_value, float& new_value, float (&w)[3], float& sq_report,
		const SlavLiveTile* tile = nullptr, const int* known_pos = nullptr) {
	const SlavGeom& g = a.g;
	const bool python = a.p.semantics != LSF_SEMANTICS_CPP;
	const bool float64 = a.p.semantics == LSF_SEMANTICS_PY_DIRECT;
#pragma unroll
	for (int c = 0; c < 3; c++) w[c] = 0.0f;
#pragma unroll
	for (int c = 0; c < D; c++) w[c] = python ? -update[c] * a.p.rate : update[c];
	float sq_before = 0.0f;
#pragma unroll
	for (int c = 0; c < D; c++) sq_before += w[c] * w[c];
	bool skip = false;
	if (a.band_union_only && slav_truncated(live_value) && slav_truncated(canonical_value)) skip = true;
	if (!skip && a.known_values_only && (float64 ? live_value == 1.0f : fabsf(live_value) == 1.0f)) skip = true;
	new_value = live_value;
	if (!skip) {
		int pos[3];
		if (known_pos != nullptr) {
#pragma unroll
			for (int ax = 0; ax < 3; ax++) pos[ax] = known_pos[ax];
		} else
			slav_coords<D>(g, idx, pos);
		int base[3] = { 0, 0, 0 };
		const float oob = a.substitute_original ? live_value : 1.0f;
		double result;
		if (float64) {
			double ratio[3] = { 0, 0, 0 };
#pragma unroll
			for (int c = 0; c < D; c++) {
				const int ax = slav_axis<D>(g, c);
				const double lookup = (double) pos[ax] + (double) w[c];
				const double fl = floor(lookup);
				base[ax] = (int) fl;
				ratio[ax] = lookup - fl;
			}
			double value[1 << D];
#pragma unroll
			for (int corner = 0; corner < (1 << D); corner++) {
				int q[3] = { 0, 0, 0 };
#pragma unroll
				for (int ax = 0; ax < D; ax++) q[ax] = base[ax] + ((corner >> ax) & 1);
				value[corner] = (double) (slav_inside<D>(g, q) ? SLAV_LIVE_TAP(a.live, q[0], (slav_index<D>(g, q))) : oob);
			}
#pragma unroll
			for (int c = D - 1; c >= 0; c--) {
				const int ax = slav_axis<D>(g, c);
#pragma unroll
				for (int corner = 0; corner < (1 << D); corner++) {
					if ((corner >> ax) & 1) continue;
					value[corner] = value[corner] * (1.0 - ratio[ax]) + value[corner | (1 << ax)] * ratio[ax];
				}
			}
			result = value[0];
			new_value = (float) result;
		} else {
			float ratio[3] = { 0.f, 0.f, 0.f };
#pragma unroll
			for (int c = 0; c < D; c++) {
				const int ax = slav_axis<D>(g, c);
				const float lookup = (float) pos[ax] + w[c];
				base[ax] = __float2int_rd(lookup);
				ratio[ax] = lookup - (float) base[ax];
			}
			float value[1 << D];
			bool staged = TILE;
			if (TILE) {
#pragma unroll
				for (int ax = 0; ax < D; ax++)
					staged = staged && base[ax] >= tile->lo[ax] && base[ax] + 1 < tile->lo[ax] + tile->ext[ax] && base[ax] >= 0
							&& base[ax] + 1 < g.n[ax];
			}
			if (TILE && staged) {
				const float* cell = tile->data
						+ ((base[0] - tile->lo[0]) * tile->ext[1] + (base[1] - tile->lo[1])) * tile->ext[2] + (base[2] - tile->lo[2]);
				__builtin_assume(__isShared(cell));
#pragma unroll
				for (int corner = 0; corner < (1 << D); corner++)
					value[corner] = cell[((corner & 1) * tile->ext[1] + ((corner >> 1) & 1)) * tile->ext[2] + ((corner >> 2) & 1)];
			} else {
#pragma unroll
				for (int corner = 0; corner < (1 << D); corner++) {
					int q[3] = { 0, 0, 0 };
#pragma unroll
					for (int ax = 0; ax < D; ax++) q[ax] = base[ax] + ((corner >> ax) & 1);
					value[corner] = slav_inside<D>(g, q) ? SLAV_LIVE_TAP(a.live, q[0], (slav_index<D>(g, q))) : oob;
				}
			}
			// interpolation along the last component's axis first (reference field_warping.tpp:126-134,187-189)
#pragma unroll
			for (int c = D - 1; c >= 0; c--) {
				const int ax = slav_axis<D>(g, c);
				const float r = ratio[ax], inverse = 1.0f - r;
#pragma unroll
				for (int corner = 0; corner < (1 << D); corner++) {
					if ((corner >> ax) & 1) continue;
					value[corner] = value[corner] * inverse + value[corner | (1 << ax)] * r;
				}
			}
			new_value = value[0];
			result = (double) new_value;
		}
		const bool snaps = float64 ? (1.0 - fabs(result) < 1e-6) : (1.0 - fabs((double) new_value) < (double) a.threshold);
		if (a.modify_warp && snaps) {
			if (float64) new_value = result > 0.0 ? 1.0f : (result < 0.0 ? -1.0f : 0.0f);
			else new_value = copysignf(1.0f, new_value);
#pragma unroll
			for (int c = 0; c < D; c++) w[c] = 0.0f;
			if (a.gradient_field != nullptr) {
#pragma unroll
				for (int c = 0; c < D; c++) a.gradient_field[c * g.N + idx] = 0.0f;
			}
		}
	}
	if (python) sq_report = fmaxf(sq_report, sq_before);
	else {
		float sq = 0.0f;
#pragma unroll
		for (int c = 0; c < D; c++) sq += w[c] * w[c];
		sq_report = fmaxf(sq_report, sq);
	}
}

template<int D>
__device__ __forceinline__ void slav_resample_at(const SlavResampleArgs& a, int idx, float& sq_report,
		const int* known_pos = nullptr) {
	const SlavGeom& g = a.g;
	float update[3] = { 0.f, 0.f, 0.f }, w[3], new_value;
#pragma unroll
	for (int c = 0; c < D; c++) update[c] = __ldg(a.update + c * g.N + idx);
	const float live_value = __ldg(a.live + idx);
	const float canonical_value = a.band_union_only ? __ldg(a.canonical + idx) : 0.0f;
	slav_resample_voxel<D>(a, idx, update, live_value, canonical_value, new_value, w, sq_report, nullptr, known_pos);
	a.new_live[idx] = new_value;
	if (a.warp != nullptr) {
#pragma unroll
		for (int c = 0; c < D; c++) a.warp[c * g.N + idx] = w[c];
	}
}

template<int D>
static __global__ void __launch_bounds__(256) k_slav_resample(SlavResampleArgs a) {
	if (a.status != nullptr && a.status[a.iteration]) return;
	const long long linear = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	float sq_report = 0.0f;
	if (linear < a.g.N) slav_resample_at<D>(a, (int) linear, sq_report);
	if (a.max_sq_bits != nullptr) block_atomic_max(sq_report, a.max_sq_bits);
}

// The same re-warp with four consecutive voxels of the last axis per thread (128-bit loads and stores; four voxels'
// loads in flight per thread). Requires n[D-1] % 4 == 0, 16-byte aligned fields and a warp output.
template<int D>
static __global__ void __launch_bounds__(256) k_slav_resample_v4(SlavResampleArgs a) {
	if (a.status != nullptr && a.status[a.iteration]) return;
	const SlavGeom& g = a.g;
	const long long group = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	float sq_report = 0.0f;
	if (group * 4 < g.N) {
		const int base = (int) (group * 4);
		float4 u4[3];
#pragma unroll
		for (int c = 0; c < D; c++) u4[c] = __ldg(reinterpret_cast<const float4*>(a.update + c * g.N + base));
		const float4 live4 = __ldg(reinterpret_cast<const float4*>(a.live + base));
		const float4 canonical4 = a.band_union_only ? __ldg(reinterpret_cast<const float4*>(a.canonical + base))
				: make_float4(0.f, 0.f, 0.f, 0.f);
		const float live_v[4] = { live4.x, live4.y, live4.z, live4.w };
		const float canonical_v[4] = { canonical4.x, canonical4.y, canonical4.z, canonical4.w };
		float out_live[4], out_w[3][4];
#pragma unroll
		for (int v = 0; v < 4; v++) {
			float update[3] = { 0.f, 0.f, 0.f }, w[3];
#pragma unroll
			for (int c = 0; c < D; c++) update[c] = v == 0 ? u4[c].x : (v == 1 ? u4[c].y : (v == 2 ? u4[c].z : u4[c].w));
			slav_resample_voxel<D>(a, base + v, update, live_v[v], canonical_v[v], out_live[v], w, sq_report);
#pragma unroll
			for (int c = 0; c < D; c++) out_w[c][v] = w[c];
		}
		*reinterpret_cast<float4*>(a.new_live + base) = make_float4(out_live[0], out_live[1], out_live[2], out_live[3]);
#pragma unroll
		for (int c = 0; c < D; c++)
			*reinterpret_cast<float4*>(a.warp + c * g.N + base) = make_float4(out_w[c][0], out_w[c][1], out_w[c][2], out_w[c][3]);
	}
	if (a.max_sq_bits != nullptr) block_atomic_max(sq_report, a.max_sq_bits);
}

// k_slav_resample_v4 with the band compacted per block (see k_slav_gradient_cpp3_band): voxels skipped by the band-union
// rule (reference field_warping.cpp:88-91) keep their live value and take the update vector as their warp directly;
// the others are listed in shared memory and re-warped by the whole block after a barrier. Requires band_union_only,
// n[D-1] % 4 == 0, 16-byte aligned fields and a warp output.
template<int D>
static __global__ void __launch_bounds__(256) k_slav_resample_band(SlavResampleArgs a) {
	if (a.status != nullptr && a.status[a.iteration]) return;
	__shared__ unsigned short band_list[1024];
	__shared__ int band_count;
	const SlavGeom& g = a.g;
	const long long block_base = (long long) blockIdx.x * 1024;
	if (threadIdx.x == 0) band_count = 0;
	__syncthreads();
	float sq_report = 0.0f;
	const long long first = block_base + threadIdx.x * 4;
	if (first < g.N) {
		const int base = (int) first;
		float4 u4[3];
#pragma unroll
		for (int c = 0; c < D; c++) u4[c] = __ldg(reinterpret_cast<const float4*>(a.update + c * g.N + base));
		const float4 live4 = __ldg(reinterpret_cast<const float4*>(a.live + base));
		const float4 canonical4 = __ldg(reinterpret_cast<const float4*>(a.canonical + base));
		const float live_v[4] = { live4.x, live4.y, live4.z, live4.w };
		const float canonical_v[4] = { canonical4.x, canonical4.y, canonical4.z, canonical4.w };
		unsigned in_band = 0;
#pragma unroll
		for (int v = 0; v < 4; v++) {
			if (!(slav_truncated(live_v[v]) && slav_truncated(canonical_v[v]))) in_band |= 1u << v;
			else {
				float sq = 0.0f;
#pragma unroll
				for (int c = 0; c < D; c++) {
					const float w = v == 0 ? u4[c].x : (v == 1 ? u4[c].y : (v == 2 ? u4[c].z : u4[c].w));
					sq += w * w;
				}
				sq_report = fmaxf(sq_report, sq);
			}
		}
		if (in_band != 0) {
			const int at = atomicAdd(&band_count, __popc(in_band));
			int k = 0;
#pragma unroll
			for (int v = 0; v < 4; v++)
				if (in_band & (1u << v)) band_list[at + k++] = (unsigned short) (threadIdx.x * 4 + v);
		}
		// skipped voxels: new live = live, warp = update (the band voxels' slots are overwritten after the barrier)
		*reinterpret_cast<float4*>(a.new_live + base) = live4;
#pragma unroll
		for (int c = 0; c < D; c++) *reinterpret_cast<float4*>(a.warp + c * g.N + base) = u4[c];
	}
	__syncthreads();
	const int count = band_count;
	for (int j = threadIdx.x; j < count; j += 256) {
		const int idx = (int) block_base + band_list[j];
		float update[3] = { 0.f, 0.f, 0.f }, w[3], new_value;
#pragma unroll
		for (int c = 0; c < D; c++) update[c] = __ldg(a.update + c * g.N + idx);
		slav_resample_voxel<D>(a, idx, update, __ldg(a.live + idx), __ldg(a.canonical + idx), new_value, w, sq_report);
		a.new_live[idx] = new_value;
#pragma unroll
		for (int c = 0; c < D; c++) a.warp[c * g.N + idx] = w[c];
	}
	if (a.max_sq_bits != nullptr) block_atomic_max(sq_report, a.max_sq_bits);
}


// termination test on the device, sticky: status[it + 1] = status[it] || finished(it + 1, max of iteration it)
// reference optimizer2d.cpp:76-82 (C++), slavcheva_optimizer2d.py:360-362 (Python)
__device__ __forceinline__ bool slav_finished(const SlavParams& p, int completed, int max_iterations, float max_warp) {
	if (p.semantics == LSF_SEMANTICS_CPP)
		return completed >= p.min_iterations && (completed >= max_iterations || max_warp < p.lower || max_warp > p.upper);
	return !(completed < p.min_iterations || (completed < max_iterations && p.lower < max_warp && max_warp < p.upper));
}

static __global__ void k_slav_decide(SlavParams p, const unsigned* max_sq_bits, int* status, int iteration,
		int max_iterations) {
	if (status[iteration]) {
		status[iteration + 1] = 1;
		return;
	}
	const float max_warp = sqrtf(__uint_as_float(max_sq_bits[iteration]));
	status[iteration + 1] = slav_finished(p, iteration + 1, max_iterations, max_warp) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------- telemetry
// reference cpp/src/telemetry/warp_delta_statistics.tpp:88-116, tsdf_difference_statistics.tpp:86-97,
// cpp/src/math/filtered_statistics.tpp:34-144. Sums are accumulated in double (the reference does the same for the
// warp lengths); maxima carry their location as the index in the reference's traversal order so that ties resolve
// to the same voxel.
struct StatsAccumulators {
	double sum;                 // pass 0: total length / total difference; pass 1: total squared deviation
	double count;               // band-union voxels (exact in double up to 2^53)
	double above;               // lengths above the minimum threshold
	unsigned long long max_key; // (float bits << 32) | (~order index): max with first-in-order tie break
	unsigned min_bits;
	unsigned pad;
};

template<int D>
__device__ __forceinline__ unsigned eigen_order_index(const SlavGeom& g, const int (&pos)[3]) {
	unsigned k = 0, scale = 1;
#pragma unroll
	for (int a = 0; a < D; a++) {
		k += (unsigned) pos[a] * scale;
		scale *= (unsigned) g.n[a];
	}
	return k;
}

// block-wide reductions of the statistics kernels (256 threads); results are valid in thread 0
struct StatsBlock {
	double sum, count, above;
	unsigned long long key;
	unsigned min_bits;
};
__device__ __forceinline__ StatsBlock stats_block_reduce(StatsBlock v) {
	__shared__ StatsBlock partial[8];
#pragma unroll
	for (int offset = 16; offset > 0; offset >>= 1) {
		v.sum += __shfl_xor_sync(0xffffffffu, v.sum, offset);
		v.count += __shfl_xor_sync(0xffffffffu, v.count, offset);
		v.above += __shfl_xor_sync(0xffffffffu, v.above, offset);
		const unsigned long long other_key = __shfl_xor_sync(0xffffffffu, v.key, offset);
		v.key = other_key > v.key ? other_key : v.key;
		v.min_bits = min(v.min_bits, __shfl_xor_sync(0xffffffffu, v.min_bits, offset));
	}
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (lane == 0) partial[warp] = v;
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int i = 1; i < (int) (blockDim.x >> 5); i++) {
			v.sum += partial[i].sum;
			v.count += partial[i].count;
			v.above += partial[i].above;
			v.key = partial[i].key > v.key ? partial[i].key : v.key;
			v.min_bits = min(v.min_bits, partial[i].min_bits);
		}
	}
	return v;
}
__device__ __forceinline__ void stats_commit(const StatsBlock& v, int pass, StatsAccumulators* acc) {
	if (threadIdx.x != 0) return;
	atomicAdd(&acc->sum, v.sum);
	if (pass == 0) {
		atomicAdd(&acc->count, v.count);
		atomicAdd(&acc->above, v.above);
		atomicMax(&acc->max_key, v.key);
		atomicMin(&acc->min_bits, v.min_bits);
	}
}

// vector field statistics over the band union; element (voxel i, component c) = field[c * component_stride + i * voxel_stride]
// (interleaved: 1, D; planes: N, 1). pass 0: sums and extrema; pass 1: squared deviation from `mean`
template<int D>
static __global__ void __launch_bounds__(256) k_warp_statistics(SlavGeom g, const float* __restrict__ field,
		long long component_stride, int voxel_stride, const float* __restrict__ live, const float* __restrict__ canonical,
		float min_threshold, int pass, float mean, StatsAccumulators* acc) {
	StatsBlock v = { 0.0, 0.0, 0.0, 0ull, 0x7f800000u };
	const float threshold_sq = min_threshold * min_threshold;
	for (long long linear = (long long) blockIdx.x * blockDim.x + threadIdx.x; linear < g.N;
			linear += (long long) gridDim.x * blockDim.x) {
		const int idx = (int) linear;
		float sq = 0.0f;
#pragma unroll
		for (int c = 0; c < D; c++) {
			const float component = field[c * component_stride + (long long) idx * voxel_stride];
			sq += component * component;
		}
		if (pass == 0) {
			int pos[3];
			slav_coords<D>(g, idx, pos);
			const unsigned long long candidate = ((unsigned long long) __float_as_uint(sq) << 32)
					| (unsigned long long) (0xffffffffu - eigen_order_index<D>(g, pos));
			if (candidate > v.key) v.key = candidate;
			v.min_bits = min(v.min_bits, __float_as_uint(sq));
		}
		if (slav_truncated(live[idx]) && slav_truncated(canonical[idx])) continue;
		const float length = sqrtf(sq);
		if (pass == 0) {
			v.sum += (double) length;
			v.count += 1.0;
			if (sq > threshold_sq) v.above += 1.0;
		} else {
			float deviation = length - mean;
			deviation = deviation * deviation;
			v.sum += (double) deviation;
		}
	}
	stats_commit(stats_block_reduce(v), pass, acc);
}

template<int D>
static __global__ void __launch_bounds__(256) k_difference_statistics(SlavGeom g, const float* __restrict__ live,
		const float* __restrict__ canonical, int pass, float mean, StatsAccumulators* acc) {
	StatsBlock v = { 0.0, 0.0, 0.0, 0ull, 0x7f800000u };
	for (long long linear = (long long) blockIdx.x * blockDim.x + threadIdx.x; linear < g.N;
			linear += (long long) gridDim.x * blockDim.x) {
		const int idx = (int) linear;
		const float d = fabsf(live[idx] - canonical[idx]);
		if (pass == 0) {
			int pos[3];
			slav_coords<D>(g, idx, pos);
			const unsigned long long candidate = ((unsigned long long) __float_as_uint(d) << 32)
					| (unsigned long long) (0xffffffffu - eigen_order_index<D>(g, pos));
			if (candidate > v.key) v.key = candidate;
			v.min_bits = min(v.min_bits, __float_as_uint(d));
			v.sum += (double) d;
		} else {
			const float deviation = d - mean;
			v.sum += (double) (deviation * deviation);
		}
	}
	stats_commit(stats_block_reduce(v), pass, acc);
}

#endif  // __CUDACC__

}  // namespace lsf
