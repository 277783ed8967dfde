/*
 * lsf_oracle_slavcheva.cpp -- CPU ORACLE (test infrastructure, NOT product code; see lsf_oracle.h) for the
 * SobolevFusion / KillingFusion ("slavcheva") optimizers, dimension-generic (2D and 3D).
 *
 * Three semantics are restated, selected by orc_slavcheva_params.semantics:
 *   ORC_SEMANTICS_CPP (0)     the reference's C++ SobolevOptimizer2d
 *                             (cpp/src/nonrigid_optimization/slavcheva/sobolev_optimizer2d.cpp:71-138, optimizer2d.cpp:76-82,
 *                              data_term.cpp:63-84, smoothing_term.cpp:26-110, cpp/src/math/convolution.cpp:69-145,
 *                              field_warping.cpp:64-136, cpp/src/math/statistics.tpp:57-100), float32 in source order.
 *                             The 3D form and the Killing / level-set terms under these semantics are a dimensional
 *                             generalisation (the reference has no 3D slavcheva optimizer): the term
 *                             formulas are the reference's Python ones (smoothing_term.py:50-100, level_set_term.py:28-64,
 *                             quirks F16 kept), the loop structure, masks, filter and resampling are the C++ ones.
 *   ORC_SEMANTICS_PY_DIRECT (1)      the reference's Python SlavchevaOptimizer2d, ComputeMethod.DIRECT
 *                             (nonrigid_opt/slavcheva/slavcheva_optimizer2d.py:238-330,332-408; field_warping.py:112-151;
 *                              math_utils/convolution.py:114-132; utils/sampling.py)
 *   ORC_SEMANTICS_PY_VECTORIZED (2)  the same class, ComputeMethod.VECTORIZED (slavcheva_optimizer2d.py:163-236), which ends
 *                             every iteration with the C++ warp_field_advanced (default flags).
 *
 * Pinning: CPP semantics against the reference's own golden vectors (cpp/tests/test_slavcheva_optimizer.cpp,
 * cpp/tests/data/test_data_slavcheva_optimizer.hpp, tests/test_slavcheva_optimizer.py, tests/test_field_warping.py);
 * the Python semantics against runs of the reference's Python class (tests/golden/make_golden.py). The 3D forms have
 * no reference counterpart: "parity unpinned" except through the degenerate-3D == 2D property (tests/).
 *
 * Conventions: see lsf_oracle.h. Component c of a vector displaces along array axis comp_axis(c):
 * 2D: c=0 (u) -> axis 1 (columns, "x"), c=1 (v) -> axis 0 (rows, "y"); 3D: c -> axis c.
 */
#include "lsf_oracle.h"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>
#include <omp.h>

namespace {

typedef std::vector<float> Field;

struct Geometry {
	int nd;
	int n[3];        // extent per array axis
	long stride[3];  // element stride per array axis (scalar field)
	long N;
	int comp_axis[3];
	Geometry(int nd_, const int* dims) : nd(nd_) {
		N = 1;
		for (int a = 0; a < 3; a++) n[a] = a < nd ? dims[a] : 1;
		for (int a = nd - 1; a >= 0; a--) {
			stride[a] = N;
			N *= n[a];
		}
		for (int a = nd; a < 3; a++) stride[a] = 0;
		if (nd == 2) {
			comp_axis[0] = 1;
			comp_axis[1] = 0;
			comp_axis[2] = 0;
		} else {
			comp_axis[0] = 0;
			comp_axis[1] = 1;
			comp_axis[2] = 2;
		}
	}
	void coords(long idx, int* p) const {
		for (int a = 0; a < nd; a++) {
			p[a] = (int) (idx / stride[a]);
			idx -= (long) p[a] * stride[a];
		}
	}
	bool inside(const int* p) const {
		for (int a = 0; a < nd; a++)
			if (p[a] < 0 || p[a] >= n[a]) return false;
		return true;
	}
	long index(const int* p) const {
		long idx = 0;
		for (int a = 0; a < nd; a++) idx += p[a] * stride[a];
		return idx;
	}
	// k-th element in the reference's traversal order (Eigen column-major: first array axis fastest)
	long eigen_order(long k) const {
		long idx = 0;
		for (int a = 0; a < nd; a++) {
			idx += (k % n[a]) * stride[a];
			k /= n[a];
		}
		return idx;
	}
};

// reference cpp/src/math/boolean_operations.hpp:37-39, utils/tsdf_set_routines.py:42-52
inline bool truncated(float v) {
	return std::fabs(v) == 1.0f;
}
inline bool both_truncated(float a, float b) {
	return truncated(a) && truncated(b);
}

// central differences, one-sided at the borders (cpp/src/math/gradients.tpp:248-283; np.gradient)
inline float central_or_one_sided(const float* f, long idx, long stride, int i, int n) {
	if (i == 0) return f[idx + stride] - f[idx];
	if (i == n - 1) return f[idx] - f[idx - stride];
	return 0.5f * (f[idx + stride] - f[idx - stride]);
}

// scalar sample with out-of-bounds -> 1 (utils/sampling.py:35-55, field_warping.tpp:29-43)
inline float sample_or_one(const Geometry& g, const float* f, const int* p) {
	return g.inside(p) ? f[g.index(p)] : 1.0f;
}

struct Terms {
	const orc_slavcheva_params* p;
	const Geometry& g;
	const float* live;
	const float* canonical;
	const float* warp;  // [N][nd] from the previous iteration
	Terms(const orc_slavcheva_params* p_, const Geometry& g_, const float* live_, const float* canonical_,
			const float* warp_) : p(p_), g(g_), live(live_), canonical(canonical_), warp(warp_) {
	}

	// ---- data term -------------------------------------------------------------------------------------
	void live_gradient(long idx, const int* pos, float* out) const {
		for (int c = 0; c < g.nd; c++) {
			const int a = g.comp_axis[c];
			out[c] = g.n[a] < 2 ? 0.0f : central_or_one_sided(live, idx, g.stride[a], pos[a], g.n[a]);
		}
		if (p->data_term_method == ORC_DATA_TERM_THRESHOLDED_FDM) {
			// nonrigid_opt/slavcheva/data_term.py:190-227
			for (int c = 0; c < g.nd; c++) {
				if (std::fabs(out[c]) > 0.5f) {
					const int a = g.comp_axis[c];
					int q[3] = { pos[0], pos[1], pos[2] };
					q[a] = pos[a] - 1;
					const float minus = sample_or_one(g, live, q);
					q[a] = pos[a] + 1;
					const float plus = sample_or_one(g, live, q);
					const float forward = plus - live[idx], backward = live[idx] - minus;
					float value = std::fabs(forward) < std::fabs(backward) ? forward : backward;
					if (std::fabs(value) > 0.5f) value = 0.0f;
					out[c] = value;
				}
			}
		}
	}
	void data_term(long idx, const int* pos, float* out) const {
		float grad[3];
		live_gradient(idx, pos, grad);
		const float diff = live[idx] - canonical[idx];
		if (p->semantics == ORC_SEMANTICS_CPP) {
			// data_term.cpp:81: scaling_factor * diff * local_live_gradient
			const float scaled = 10.0f * diff;
			for (int c = 0; c < g.nd; c++) out[c] = scaled * grad[c];
		} else {
			// data_term.py:169-187,334-351: diff * live_local_gradient * scaling_factor
			for (int c = 0; c < g.nd; c++) out[c] = (diff * grad[c]) * 10.0f;
		}
	}

	// ---- smoothing terms -------------------------------------------------------------------------------
	// warp sample with out-of-bounds -> centre value (utils/sampling.py:84-88)
	inline float warp_or_centre(const int* q, int c, long centre_idx) const {
		return g.inside(q) ? warp[g.index(q) * g.nd + c] : warp[centre_idx * g.nd + c];
	}
	// C++ Tikhonov: smoothing_term.cpp:43-108 -- per axis, array axis 0 assigned, then 1 (then 2) added
	void tikhonov_cpp(long idx, const int* pos, float* out) const {
		for (int c = 0; c < g.nd; c++) {
			float total = 0.0f;
			for (int a = 0; a < g.nd; a++) {
				const long s = g.stride[a] * g.nd;
				const float* w = warp + idx * g.nd + c;
				const int i = pos[a], n = g.n[a];
				float term;
				if (n < 2) term = 0.0f;
				else if (i == 0) term = -w[s] + w[0];
				else if (i == n - 1) term = -w[-s] + w[0];
				else term = (-w[s] + 2.0f * w[0]) - w[-s];
				if (a == 0) total = term;
				else total += term;
			}
			out[c] = total;
		}
	}
	// Python Tikhonov: smoothing_term.py:103-139 (copy_if_zero=False): -(x+1 + y+1 - 4w + x-1 + y-1), neighbours outside
	// the field replaced by the centre value. 3D: -(sum over component order of the +1 neighbours - 2*nd*w + the -1s).
	void tikhonov_py(long idx, const int* pos, float* out) const {
		for (int c = 0; c < g.nd; c++) {
			float acc = 0.0f;
			int q[3] = { pos[0], pos[1], pos[2] };
			for (int k = 0; k < g.nd; k++) {  // + neighbours in component order (x, y[, z])
				const int a = g.comp_axis[k];
				q[a] = pos[a] + 1;
				const float v = warp_or_centre(q, c, idx);
				q[a] = pos[a];
				acc = k == 0 ? v : acc + v;
			}
			acc = acc - (2.0f * g.nd) * warp[idx * g.nd + c];
			for (int k = 0; k < g.nd; k++) {
				const int a = g.comp_axis[k];
				q[a] = pos[a] - 1;
				acc = acc + warp_or_centre(q, c, idx);
				q[a] = pos[a];
			}
			out[c] = -acc;
		}
	}
	// Killing: smoothing_term.py:50-100 (copy_if_zero=False), quirks F16 kept:
	//   second difference along the component-0 axis ("x") is x+1 - 2w + x-1; along every other axis it uses the +1
	//   neighbour twice; result[a] = -2(1+lambda)*xx[a] + sum_other yy[a] + lambda * sum_{b != a} mixed(a,b)[b]
	void killing(long idx, const int* pos, float* out) const {
		const float lambda = p->isomorphic_enforcement_factor;
		const float c0 = (float) (-2.0 * (1.0 + (double) lambda));
		const int ax0 = g.comp_axis[0];
		for (int a = 0; a < g.nd; a++) {
			int q[3] = { pos[0], pos[1], pos[2] };
			const float w = warp[idx * g.nd + a];
			q[ax0] = pos[ax0] + 1;
			const float xp = warp_or_centre(q, a, idx);
			q[ax0] = pos[ax0] - 1;
			const float xm = warp_or_centre(q, a, idx);
			q[ax0] = pos[ax0];
			float acc = c0 * ((xp - 2.0f * w) + xm);
			for (int k = 1; k < g.nd; k++) {
				const int ax = g.comp_axis[k];
				q[ax] = pos[ax] + 1;
				const float yp = warp_or_centre(q, a, idx);
				q[ax] = pos[ax];
				acc = acc + ((yp - 2.0f * w) + yp);
			}
			for (int b = 0; b < g.nd; b++) {
				if (b == a) continue;
				// mixed derivative of component b along the axes of components min(a,b) ("first") and max(a,b)
				const int first = g.comp_axis[std::min(a, b)], second = g.comp_axis[std::max(a, b)];
				float v[4];
				int k = 0;
				for (int s1 = 1; s1 >= -1; s1 -= 2)
					for (int s2 = 1; s2 >= -1; s2 -= 2) {
						q[first] = pos[first] + s1;
						q[second] = pos[second] + s2;
						v[k++] = warp_or_centre(q, b, idx);
					}
				q[first] = pos[first];
				q[second] = pos[second];
				// (x+1,y+1) - (x+1,y-1) - (x-1,y+1) + (x-1,y-1), / 4
				const float mixed = (((v[0] - v[1]) - v[2]) + v[3]) / 4.0f;
				acc = acc + lambda * mixed;
			}
			out[a] = acc;
		}
	}

	// ---- level-set term: level_set_term.py:28-64 (out-of-bounds live -> 1; quirk: both second differences use +1 twice)
	void level_set(long idx, const int* pos, float* out) const {
		const float centre = live[idx];
		float grad[3], hessian[3][3];
		int q[3] = { pos[0], pos[1], pos[2] };
		float plus[3];
		for (int c = 0; c < g.nd; c++) {
			const int a = g.comp_axis[c];
			q[a] = pos[a] + 1;
			plus[c] = sample_or_one(g, live, q);
			q[a] = pos[a] - 1;
			const float minus = sample_or_one(g, live, q);
			q[a] = pos[a];
			grad[c] = (0.5f * (plus[c] - minus)) * 10.0f;
			hessian[c][c] = ((plus[c] - 2.0f * centre) + plus[c]) * 10.0f;
		}
		for (int c1 = 0; c1 < g.nd; c1++)
			for (int c2 = c1 + 1; c2 < g.nd; c2++) {
				const int a1 = g.comp_axis[c1], a2 = g.comp_axis[c2];
				float v[4];
				int k = 0;
				// (x+1,y+1) - (x-1,y+1) - (x+1,y-1) + (x-1,y-1)
				for (int s2 = 1; s2 >= -1; s2 -= 2)
					for (int s1 = 1; s1 >= -1; s1 -= 2) {
						q[a1] = pos[a1] + s1;
						q[a2] = pos[a2] + s2;
						v[k++] = sample_or_one(g, live, q);
					}
				q[a1] = pos[a1];
				q[a2] = pos[a2];
				const float mixed = (0.25f * (((v[0] - v[1]) - v[2]) + v[3])) * 10.0f;
				hessian[c1][c2] = hessian[c2][c1] = mixed;
			}
		float sq = 0.0f;
		for (int c = 0; c < g.nd; c++) sq += grad[c] * grad[c];
		const float length = std::sqrt(sq);
		const float factor = (1.0f - length) / (length + 1e-5f);
		for (int a = 0; a < g.nd; a++) {
			float acc = 0.0f;
			for (int b = 0; b < g.nd; b++) acc = b == 0 ? hessian[a][0] * grad[0] : acc + hessian[a][b] * grad[b];
			out[a] = factor * acc;
		}
	}
};

// one pass of the separable filter along array axis `axis` of an interleaved vector field
// (cpp/src/math/convolution.cpp:50-67: taps i-r..i+r ascending from 0.0f, zero padded, kernel flipped)
void convolve_axis(const Geometry& g, const float* in, float* out, int axis, const float* kernel, int K,
		int zero_rule, const unsigned char* py_zero_mask) {
	const int r = K / 2;
	const int D = g.nd;
	const long s = g.stride[axis] * D;
	const int n = g.n[axis];
#pragma omp parallel for
	for (long idx = 0; idx < g.N; idx++) {
		int pos[3];
		g.coords(idx, pos);
		const int i = pos[axis];
		const float* centre = in + idx * D;
		if (zero_rule == 1) {  // C++: pass-input vector exactly zero -> zero output (convolution.cpp:23-47,69-145)
			bool all_zero = true;
			for (int c = 0; c < D; c++) all_zero = all_zero && centre[c] == 0.0f;
			if (all_zero) {
				for (int c = 0; c < D; c++) out[idx * D + c] = 0.0f;
				continue;
			}
		}
		for (int c = 0; c < D; c++) {
			float acc = 0.0f;
			for (int j = 0; j < K; j++) {
				const int src = i - r + j;
				const float value = (src >= 0 && src < n) ? centre[(long) (j - r) * s + c] : 0.0f;
				acc += value * kernel[K - 1 - j];
			}
			// Python: |component of the ORIGINAL field| < 1e-6 -> zero, per component (math_utils/convolution.py:114-127)
			if (zero_rule == 2 && py_zero_mask[idx * D + c]) acc = 0.0f;
			out[idx * D + c] = acc;
		}
	}
}

void convolve_preserve_zeros(const Geometry& g, Field& field, Field& scratch, const float* kernel, int K, int semantics) {
	const int D = g.nd;
	scratch.resize(field.size());
	std::vector<unsigned char> mask;
	int zero_rule = 1;
	if (semantics != ORC_SEMANTICS_CPP) {
		zero_rule = 2;
		mask.resize(field.size());
		for (size_t i = 0; i < field.size(); i++) mask[i] = std::fabs(field[i]) < 1e-6f;
	}
	// pass order: array axis 0 first (2D: rows/y then columns/x, convolution.cpp:69-145; 3D: axes 0,1,2, :221-332)
	float* a = field.data();
	float* b = scratch.data();
	for (int axis = 0; axis < D; axis++) {
		convolve_axis(g, a, b, axis, kernel, K, zero_rule, mask.data());
		std::swap(a, b);
	}
	if (a != field.data()) std::memcpy(field.data(), a, field.size() * sizeof(float));
}

// masked resample of the live field, truncation snap, warp zeroing
// (field_warping.cpp:64-136; Python field_warping.py:112-151 interpolates in float64)
void warp_advanced(const Geometry& g, const float* live, const float* canonical, float* warp, float* gradient_field,
		float* new_live, bool band_union_only, bool known_values_only, bool substitute_original, float threshold,
		bool modify_warp, bool python_float64) {
	const int D = g.nd;
#pragma omp parallel for
	for (long idx = 0; idx < g.N; idx++) {
		const float live_value = live[idx];
		if (band_union_only && both_truncated(live_value, canonical[idx])) {
			new_live[idx] = live_value;
			continue;
		}
		if (known_values_only) {
			if (python_float64 ? live_value == 1.0f : std::fabs(live_value) == 1.0f) {
				new_live[idx] = live_value;
				continue;
			}
		}
		int pos[3];
		g.coords(idx, pos);
		int base[3] = { 0, 0, 0 };
		double ratio[3] = { 0, 0, 0 };
		for (int c = 0; c < D; c++) {
			const int a = g.comp_axis[c];
			if (python_float64) {
				const double lookup = (double) pos[a] + (double) warp[idx * D + c];
				const double fl = std::floor(lookup);
				base[a] = (int) fl;
				ratio[a] = lookup - fl;
			} else {
				const float lookup = (float) pos[a] + warp[idx * D + c];
				base[a] = (int) std::floor(lookup);
				ratio[a] = (double) (lookup - (float) base[a]);
			}
		}
		// corner values, index bit a = +1 along array axis a
		double value[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
		const float oob = substitute_original ? live_value : 1.0f;
		for (int corner = 0; corner < (1 << D); corner++) {
			int q[3] = { 0, 0, 0 };
			for (int a = 0; a < D; a++) q[a] = base[a] + ((corner >> a) & 1);
			value[corner] = g.inside(q) ? live[g.index(q)] : oob;
		}
		// interpolation order: along the LAST component's axis first (3D: z, y, x; 2D: y then x), F9
		double result;
		{
			double current[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
			int count = 1 << D;
			for (int i = 0; i < count; i++) current[i] = value[i];
			unsigned alive = (1u << D) - 1;  // axes still un-interpolated
			for (int c = D - 1; c >= 0; c--) {
				const int a = g.comp_axis[c];
				for (int corner = 0; corner < (1 << D); corner++) {
					if ((corner & ~alive) != 0 || ((corner >> a) & 1)) continue;
					const double lo = current[corner], hi = current[corner | (1 << a)];
					if (python_float64) current[corner] = lo * (1.0 - ratio[a]) + hi * ratio[a];
					else {
						const float r = (float) ratio[a], q1 = 1.0f - r;
						current[corner] = (double) ((float) lo * q1 + (float) hi * r);
					}
				}
				alive &= ~(1u << a);
			}
			result = current[0];
		}
		float new_value = (float) result;
		const bool snaps = python_float64 ? (1.0 - std::fabs(result) < 1e-6) : (1.0 - std::fabs((double) new_value)
				< (double) threshold);
		if (modify_warp && snaps) {
			if (python_float64) new_value = result > 0 ? 1.0f : (result < 0 ? -1.0f : 0.0f);  // np.sign
			else new_value = std::copysign(1.0f, new_value);
			for (int c = 0; c < D; c++) warp[idx * D + c] = 0.0f;
			if (gradient_field) for (int c = 0; c < D; c++) gradient_field[idx * D + c] = 0.0f;
		}
		new_live[idx] = new_value;
	}
}

float max_norm(const float* v, int D, long n, bool python) {
	if (python) {  // np.linalg.norm of each float32 vector, max (slavcheva_optimizer2d.py:212-215,309-318)
		float best = 0.0f;
		for (long i = 0; i < n; i++) {
			float sq = 0.0f;
			for (int c = 0; c < D; c++) sq += v[i * D + c] * v[i * D + c];
			const float length = std::sqrt(sq);
			if (length > best) best = length;
		}
		return best;
	}
	float max_sq = 0.0f;
#pragma omp parallel for reduction(max:max_sq)
	for (long i = 0; i < n; i++) {
		float sq = 0.0f;
		for (int c = 0; c < D; c++) sq += v[i * D + c] * v[i * D + c];
		if (sq > max_sq) max_sq = sq;
	}
	return std::sqrt(max_sq);
}

struct State {
	Field live, warp, gradient_field, scratch, new_live;
};

float slavcheva_iteration(const orc_slavcheva_params* p, const Geometry& g, const float* canonical, State& s) {
	const int D = g.nd;
	const Terms terms(p, g, s.live.data(), canonical, s.warp.data());
	const bool cpp = p->semantics == ORC_SEMANTICS_CPP;
	const bool direct = p->semantics == ORC_SEMANTICS_PY_DIRECT;
	const bool level_set_on = p->level_set_term_enabled && p->semantics != ORC_SEMANTICS_PY_VECTORIZED;
	Field fresh;
	// Python DIRECT keeps stale entries of the persistent gradient field where the band-union test skips a voxel
	// (slavcheva_optimizer2d.py:261-262,300,344); the other semantics start every iteration from zeros
	Field& gradient = s.gradient_field;
	if (!direct) gradient.assign((size_t) g.N * D, 0.0f);
	float* out = gradient.data();
#pragma omp parallel for
	for (long idx = 0; idx < g.N; idx++) {
		int pos[3] = { 0, 0, 0 };
		g.coords(idx, pos);
		const float live_value = s.live[idx];
		const bool outside = both_truncated(live_value, canonical[idx]);
		float data[3], smooth[3], ls[3] = { 0.f, 0.f, 0.f };
		float* o = out + idx * D;
		if (cpp) {
			if (outside) {
				for (int c = 0; c < D; c++) o[c] = (0.0f + 0.0f * p->smoothing_term_weight) * -p->gradient_descent_rate;
				continue;
			}
			terms.data_term(idx, pos, data);
			if (p->smoothing_term_method == ORC_SMOOTHING_KILLING) terms.killing(idx, pos, smooth);
			else terms.tikhonov_cpp(idx, pos, smooth);
			const bool ls_here = level_set_on && !truncated(live_value);
			if (ls_here) terms.level_set(idx, pos, ls);
			for (int c = 0; c < D; c++) {
				// sobolev_optimizer2d.cpp:131-132: (data + smoothing * weight) * -rate; data weight / level-set are
				// the generalisation (weight 1 and no level-set term reproduce the C++ expression bit for bit)
				float total = data[c] * p->data_term_weight;
				if (ls_here) total = total + ls[c] * p->level_set_term_weight;
				total = total + smooth[c] * p->smoothing_term_weight;
				o[c] = total * -p->gradient_descent_rate;
			}
		} else if (direct) {
			if (outside) continue;  // stale entry stays
			terms.data_term(idx, pos, data);
			float total[3];
			for (int c = 0; c < D; c++) total[c] = 0.0f + p->data_term_weight * data[c];
			if (level_set_on && !truncated(live_value)) {
				terms.level_set(idx, pos, ls);
				for (int c = 0; c < D; c++) total[c] = total[c] + p->level_set_term_weight * ls[c];
			}
			if (p->smoothing_term_method == ORC_SMOOTHING_KILLING) terms.killing(idx, pos, smooth);
			else terms.tikhonov_py(idx, pos, smooth);
			for (int c = 0; c < D; c++) o[c] = total[c] + p->smoothing_term_weight * smooth[c];
		} else {  // vectorized: np.gradient data term, -scipy.ndimage.laplace (mode reflect == replicate for a 3-pt stencil)
			terms.data_term(idx, pos, data);
			terms.tikhonov_cpp(idx, pos, smooth);  // same stencil: sum over axes of -(next - 2 cur + prev) with replication
			for (int c = 0; c < D; c++) {
				const float d = outside ? 0.0f : data[c];
				const float value = p->data_term_weight * d + p->smoothing_term_weight * smooth[c];
				o[c] = outside ? 0.0f : value;
			}
		}
	}
	if (p->sobolev_smoothing_enabled && p->kernel && p->kernel_size > 0)
		convolve_preserve_zeros(g, gradient, s.scratch, p->kernel, p->kernel_size, p->semantics);
	s.new_live.resize((size_t) g.N);
	float max_warp;
	if (cpp) {
		// the C++ optimizer scaled by -rate before the filter; `gradient` already IS the warp update
		s.warp = gradient;
		warp_advanced(g, s.live.data(), canonical, s.warp.data(), nullptr, s.new_live.data(), true, false, false, 1e-6f,
				true, false);
		max_warp = max_norm(s.warp.data(), D, g.N, false);
	} else {
		for (size_t i = 0; i < gradient.size(); i++) s.warp[i] = -gradient[i] * p->gradient_descent_rate;
		max_warp = max_norm(s.warp.data(), D, g.N, true);
		if (direct) warp_advanced(g, s.live.data(), canonical, s.warp.data(), gradient.data(), s.new_live.data(), false,
				false, false, 1e-6f, true, true);
		else warp_advanced(g, s.live.data(), canonical, s.warp.data(), nullptr, s.new_live.data(), false, false, false,
				1e-6f, true, false);
	}
	s.live.swap(s.new_live);
	return max_warp;
}

}  // namespace

namespace {

// Energy aggregates of the reference's Python optimizer for the state an iteration starts from: what
// SlavchevaOptimizer2d.total_data_energy / total_smoothing_energy / total_level_set_energy hold after the gradient pass
// (appended to OptimizationLog every iteration, slavcheva_optimizer2d.py:370-374). 2D, Python semantics only:
//   DIRECT     slavcheva_optimizer2d.py:236-300 with the local energies of data_term.py:185,225 (0.5 diff^2),
//              level_set_term.py:63 (0.5 (|10 grad live| - 1)^2), smoothing_term.py:97-98 (Killing: J.J + lambda J^T.J) and
//              :137-138 (Tikhonov: 0.5 (|w_x|^2 + |w_y|^2)), summed over the band union, each times its weight;
//   VECTORIZED :163-175 with data_term.py:352-358 and smoothing_term.py:162-177 (np.gradient of the warp components).
// Evaluated in double from the float32 fields (the reference mixes float32 locals and Python floats; compared at 1e-5).
void slavcheva_energies(const orc_slavcheva_params* p, const Geometry& g, const float* live, const float* canonical,
		const float* warp, double* out) {
	out[0] = out[1] = out[2] = 0.0;
	if (g.nd != 2 || p->semantics == ORC_SEMANTICS_CPP) return;
	const Terms terms(p, g, live, canonical, warp);
	const int ax_x = g.comp_axis[0], ax_y = g.comp_axis[1];
	double data = 0.0, smoothing = 0.0, level_set = 0.0;
	for (long idx = 0; idx < g.N; idx++) {
		if (both_truncated(live[idx], canonical[idx])) continue;
		int pos[3] = { 0, 0, 0 };
		g.coords(idx, pos);
		const double diff = (double) live[idx] - (double) canonical[idx];
		data += 0.5 * diff * diff;
		if (p->semantics == ORC_SEMANTICS_PY_DIRECT) {
			if (p->level_set_term_enabled && !truncated(live[idx])) {
				double grad[2];
				for (int c = 0; c < 2; c++) {
					int q[3] = { pos[0], pos[1], pos[2] };
					const int a = g.comp_axis[c];
					q[a] = pos[a] + 1;
					const double plus = sample_or_one(g, live, q);
					q[a] = pos[a] - 1;
					const double minus = sample_or_one(g, live, q);
					grad[c] = 0.5 * (plus - minus) * 10.0;
				}
				const double length = std::sqrt(grad[0] * grad[0] + grad[1] * grad[1]);
				level_set += 0.5 * (length - 1.0) * (length - 1.0);
			}
			double wx[2], wy[2];  // d(u, v)/dx, d(u, v)/dy; neighbours outside the field = the centre value
			for (int c = 0; c < 2; c++) {
				int q[3] = { pos[0], pos[1], pos[2] };
				q[ax_x] = pos[ax_x] + 1;
				const double xp = terms.warp_or_centre(q, c, idx);
				q[ax_x] = pos[ax_x] - 1;
				const double xm = terms.warp_or_centre(q, c, idx);
				q[ax_x] = pos[ax_x];
				q[ax_y] = pos[ax_y] + 1;
				const double yp = terms.warp_or_centre(q, c, idx);
				q[ax_y] = pos[ax_y] - 1;
				const double ym = terms.warp_or_centre(q, c, idx);
				wx[c] = 0.5 * (xp - xm);
				wy[c] = 0.5 * (yp - ym);
			}
			if (p->smoothing_term_method == ORC_SMOOTHING_KILLING) {
				const double jj = wx[0] * wx[0] + wx[1] * wx[1] + wy[0] * wy[0] + wy[1] * wy[1];
				const double jtj = wx[0] * wx[0] + wy[0] * wx[1] + wx[1] * wy[0] + wy[1] * wy[1];
				smoothing += jj + (double) p->isomorphic_enforcement_factor * jtj;
			} else {
				smoothing += 0.5 * ((wx[0] * wx[0] + wx[1] * wx[1]) + (wy[0] * wy[0] + wy[1] * wy[1]));
			}
		} else {
			double aggregate = 0.0;
			for (int c = 0; c < 2; c++)
				for (int a = 0; a < 2; a++) {
					const long s = g.stride[a] * 2;
					const float* w = warp + idx * 2 + c;
					const int i = pos[a], n = g.n[a];
					double d;
					if (i == 0) d = (double) w[s] - (double) w[0];
					else if (i == n - 1) d = (double) w[0] - (double) w[-s];
					else d = 0.5 * ((double) w[s] - (double) w[-s]);
					aggregate += d * d;
				}
			smoothing += 0.5 * aggregate;
		}
	}
	out[0] = data * (double) p->data_term_weight;
	out[1] = smoothing * (double) p->smoothing_term_weight;
	out[2] = level_set * (double) p->level_set_term_weight;
}

}  // namespace

extern "C" {

int orc_slavcheva_optimize_energies(const orc_slavcheva_params* p, const float* live, const float* canonical, int nd,
		const int* dims, float* live_out, float* warp_out, int* iteration_count, float* max_warps, int max_warps_capacity,
		orc_iteration_dump* dump, double* energies, int energies_capacity);

int orc_slavcheva_optimize(const orc_slavcheva_params* p, const float* live, const float* canonical, int nd,
		const int* dims, float* live_out, float* warp_out, int* iteration_count, float* max_warps, int max_warps_capacity,
		orc_iteration_dump* dump) {
	return orc_slavcheva_optimize_energies(p, live, canonical, nd, dims, live_out, warp_out, iteration_count, max_warps,
			max_warps_capacity, dump, nullptr, 0);
}

/* the same; energies [energies_capacity][3] (or NULL) receives {data, smoothing, level set} energy of every iteration */
int orc_slavcheva_optimize_energies(const orc_slavcheva_params* p, const float* live, const float* canonical, int nd,
		const int* dims, float* live_out, float* warp_out, int* iteration_count, float* max_warps, int max_warps_capacity,
		orc_iteration_dump* dump, double* energies, int energies_capacity) {
	if (nd != 2 && nd != 3) return -1;
	for (int a = 0; a < nd; a++)
		if (dims[a] < 2) return -2;
	if (nd == 2 && dims[0] != dims[1]) return -3;  // the reference's 2D code is only correct for square fields
	const Geometry g(nd, dims);
	State s;
	s.live.assign(live, live + g.N);
	s.warp.assign((size_t) g.N * nd, 0.0f);
	s.gradient_field.assign((size_t) g.N * nd, 0.0f);
	int iteration = 0;
	if (dump) dump->count = 0;
	const float lower = p->maximum_warp_length_lower_threshold, upper = p->maximum_warp_length_upper_threshold;
	float max_warp;
	auto finished = [&]() {
		if (p->semantics == ORC_SEMANTICS_CPP)  // optimizer2d.cpp:76-82
			return iteration >= p->min_iterations && (iteration >= p->max_iterations || max_warp < lower || max_warp > upper);
		// slavcheva_optimizer2d.py:360-362
		return !(iteration < p->min_iterations || (iteration < p->max_iterations && lower < max_warp && max_warp < upper));
	};
	max_warp = p->semantics == ORC_SEMANTICS_CPP ? upper - 1.0f : INFINITY;  // sobolev_optimizer2d.cpp:77; .py:339
	while (!finished()) {
		if (energies && iteration < energies_capacity)
			slavcheva_energies(p, g, s.live.data(), canonical, s.warp.data(), energies + 3 * (size_t) iteration);
		max_warp = slavcheva_iteration(p, g, canonical, s);
		if (max_warps && iteration < max_warps_capacity) max_warps[iteration] = max_warp;
		if (dump && dump->buffer && iteration < dump->max_iterations) {
			std::memcpy(dump->buffer + (size_t) iteration * g.N * nd, s.warp.data(), (size_t) g.N * nd * sizeof(float));
			dump->count = iteration + 1;
		}
		iteration++;
	}
	if (live_out) std::memcpy(live_out, s.live.data(), (size_t) g.N * sizeof(float));
	if (warp_out) std::memcpy(warp_out, s.warp.data(), (size_t) g.N * nd * sizeof(float));
	if (iteration_count) *iteration_count = iteration;
	return 0;
}

/* single steps, exposed so the golden vectors of the reference's unit tests can pin them */
void orc_slavcheva_energies(const orc_slavcheva_params* p, const float* live, const float* canonical, const float* warp,
		int nd, const int* dims, double* out) {
	const Geometry g(nd, dims);
	slavcheva_energies(p, g, live, canonical, warp, out);
}

void orc_slavcheva_data_term(const orc_slavcheva_params* p, const float* live, const float* canonical, int nd,
		const int* dims, int band_union_only, float* out) {
	const Geometry g(nd, dims);
	const Terms terms(p, g, live, canonical, nullptr);
	for (long idx = 0; idx < g.N; idx++) {
		int pos[3] = { 0, 0, 0 };
		g.coords(idx, pos);
		if (band_union_only && both_truncated(live[idx], canonical[idx])) {
			for (int c = 0; c < nd; c++) out[idx * nd + c] = 0.0f;
			continue;
		}
		terms.data_term(idx, pos, out + idx * nd);
	}
}

void orc_slavcheva_smoothing_term(const orc_slavcheva_params* p, const float* warp, const float* live,
		const float* canonical, int nd, const int* dims, int band_union_only, float* out) {
	const Geometry g(nd, dims);
	const Terms terms(p, g, live, canonical, warp);
	for (long idx = 0; idx < g.N; idx++) {
		int pos[3] = { 0, 0, 0 };
		g.coords(idx, pos);
		if (band_union_only && both_truncated(live[idx], canonical[idx])) {
			for (int c = 0; c < nd; c++) out[idx * nd + c] = 0.0f;
			continue;
		}
		if (p->smoothing_term_method == ORC_SMOOTHING_KILLING) terms.killing(idx, pos, out + idx * nd);
		else if (p->semantics == ORC_SEMANTICS_PY_DIRECT) terms.tikhonov_py(idx, pos, out + idx * nd);
		else terms.tikhonov_cpp(idx, pos, out + idx * nd);
	}
}

void orc_slavcheva_level_set_term(const orc_slavcheva_params* p, const float* live, int nd, const int* dims, float* out) {
	const Geometry g(nd, dims);
	const Terms terms(p, g, live, nullptr, nullptr);
	for (long idx = 0; idx < g.N; idx++) {
		int pos[3] = { 0, 0, 0 };
		g.coords(idx, pos);
		terms.level_set(idx, pos, out + idx * nd);
	}
}

void orc_warp_advanced(const float* live, const float* canonical, float* warp, int nd, const int* dims,
		int band_union_only, int known_values_only, int substitute_original, float truncation_float_threshold,
		int modify_warp, float* new_live) {
	const Geometry g(nd, dims);
	warp_advanced(g, live, canonical, warp, nullptr, new_live, band_union_only != 0, known_values_only != 0,
			substitute_original != 0, truncation_float_threshold, modify_warp != 0, false);
}

/* telemetry: cpp/src/telemetry/warp_delta_statistics.tpp:88-116, tsdf_difference_statistics.tpp:86-97,
 * cpp/src/math/filtered_statistics.tpp:34-144, statistics.tpp */
void orc_warp_delta_statistics(const float* warp, const float* canonical, const float* live, int nd, const int* dims,
		float min_threshold, float max_threshold, orc_warp_delta_statistics_t* out) {
	const Geometry g(nd, dims);
	double total_length = 0.0;
	long count = 0, above = 0;
	float max_sq = 0.0f, min_sq = FLT_MAX;
	long max_at = 0;
	const float threshold_sq = min_threshold * min_threshold;
	for (long k = 0; k < g.N; k++) {
		const long idx = g.eigen_order(k);  // ties resolve to the first element in the reference's traversal order
		float sq = 0.0f;
		for (int c = 0; c < nd; c++) sq += warp[idx * nd + c] * warp[idx * nd + c];
		if (sq > max_sq) {
			max_sq = sq;
			max_at = idx;
		}
		if (sq < min_sq) min_sq = sq;
		if (both_truncated(live[idx], canonical[idx])) continue;
		total_length += std::sqrt(sq);
		if (sq > threshold_sq) above++;
		count++;
	}
	const float mean = (float) (total_length / (double) count);
	double deviation = 0.0;
	for (long idx = 0; idx < g.N; idx++) {
		if (both_truncated(live[idx], canonical[idx])) continue;
		float sq = 0.0f;
		for (int c = 0; c < nd; c++) sq += warp[idx * nd + c] * warp[idx * nd + c];
		float local = std::sqrt(sq) - mean;
		local = local * local;
		deviation += local;
	}
	out->ratio_above_min_threshold = (float) ((double) above / (double) count);
	out->length_min = std::sqrt(min_sq);
	out->length_max = std::sqrt(max_sq);
	out->length_mean = mean;
	out->length_standard_deviation = (float) std::sqrt(deviation / (double) count);
	int pos[3] = { 0, 0, 0 };
	g.coords(max_at, pos);
	// reported in component order: (x, y[, z]) = position along comp_axis(0), comp_axis(1), ...
	for (int c = 0; c < 3; c++) out->longest_warp_location[c] = c < nd ? pos[g.comp_axis[c]] : 0;
	out->is_largest_below_min_threshold = out->length_max < min_threshold;
	out->is_largest_above_max_threshold = out->length_max > max_threshold;
}

void orc_tsdf_difference_statistics(const float* canonical, const float* live, int nd, const int* dims,
		orc_tsdf_difference_statistics_t* out) {
	const Geometry g(nd, dims);
	float dmin = FLT_MAX, dmax = -FLT_MAX;
	long at = 0;
	double sum = 0.0;
	for (long k = 0; k < g.N; k++) {
		const long idx = g.eigen_order(k);
		const float d = std::fabs(live[idx] - canonical[idx]);
		if (d < dmin) dmin = d;
		if (d > dmax) {
			dmax = d;
			at = idx;
		}
		sum += d;
	}
	const float mean = (float) (sum / (double) g.N);
	double dev = 0.0;
	for (long idx = 0; idx < g.N; idx++) {
		const float d = std::fabs(live[idx] - canonical[idx]) - mean;
		dev += (double) (d * d);
	}
	out->difference_min = dmin;
	out->difference_max = dmax;
	out->difference_mean = mean;
	out->difference_standard_deviation = (float) std::sqrt(dev / (double) g.N);
	int pos[3] = { 0, 0, 0 };
	g.coords(at, pos);
	for (int c = 0; c < 3; c++) out->biggest_difference_location[c] = c < nd ? pos[g.comp_axis[c]] : 0;
}

}  // extern "C"
