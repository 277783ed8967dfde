/*
 * lsf_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code; see lsf_oracle.h).
 *
 * A from-scratch restatement, in plain C++17 + OpenMP over flat float arrays with numpy index
 * semantics, of the reference's warp-field optimisation path. No Eigen, no Boost. Every function
 * cites the reference file:line whose arithmetic (operation order, float32, no FMA) it follows.
 *
 * Pinning status: every 2D function here is pinned against the reference's own golden vectors
 * (tests/golden/ npz files, made by tests/golden/make_golden.py from /root/reference/tests/test_data and
 * from running the reference's Python optimizers). 3D functions share the per-axis arithmetic with
 * the pinned 2D ones; the reference holds no 3D optimizer fixtures, those that
 * exist (3D gradient, laplacian, convolution, resampling literals) are pinned.
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

inline bool is_power_of_two(int v) {
	return v > 0 && (v & (v - 1)) == 0;
}

// ------------------------------------------------------------------------------------------------
// Trilinear / bilinear gather. Reference: cpp/src/nonrigid_optimization/field_warping.tpp:68-140 (3D),
// :145-194 (2D); out-of-bounds taps take `oob` (1.0 for `warp`, the replacement for
// `warp_with_replacement`), :29-64.
// ------------------------------------------------------------------------------------------------
void warp2d_generic(const float* f, int C, const float* w, int H, int W, float oob, float* out) {
	const long n = (long) H * W;
#pragma omp parallel for
	for (long idx = 0; idx < n; idx++) {
		const int y = (int) (idx / W), x = (int) (idx % W);
		const float lookup_x = (float) x + w[2 * idx];      // u displaces along columns
		const float lookup_y = (float) y + w[2 * idx + 1];  // v displaces along rows
		const int base_x = (int) std::floor(lookup_x);
		const int base_y = (int) std::floor(lookup_y);
		const float ratio_x = lookup_x - (float) base_x;
		const float ratio_y = lookup_y - (float) base_y;
		const float inv_x = 1.0f - ratio_x;
		const float inv_y = 1.0f - ratio_y;
		const bool x0 = base_x >= 0 && base_x < W, x1 = base_x + 1 >= 0 && base_x + 1 < W;
		const bool y0 = base_y >= 0 && base_y < H, y1 = base_y + 1 >= 0 && base_y + 1 < H;
		for (int c = 0; c < C; c++) {
			const float v00 = (x0 && y0) ? f[((long) base_y * W + base_x) * C + c] : oob;
			const float v01 = (x0 && y1) ? f[((long) (base_y + 1) * W + base_x) * C + c] : oob;
			const float v10 = (x1 && y0) ? f[((long) base_y * W + base_x + 1) * C + c] : oob;
			const float v11 = (x1 && y1) ? f[((long) (base_y + 1) * W + base_x + 1) * C + c] : oob;
			// y (rows) first, then x -- field_warping.tpp:187-189
			const float i0 = v00 * inv_y + v01 * ratio_y;
			const float i1 = v10 * inv_y + v11 * ratio_y;
			out[idx * C + c] = i0 * inv_x + i1 * ratio_x;
		}
	}
}

void warp3d_generic(const float* f, int C, const float* w, int X, int Y, int Z, float oob, float* out) {
	const long n = (long) X * Y * Z;
	const long sY = Z, sX = (long) Y * Z;
#pragma omp parallel for
	for (long idx = 0; idx < n; idx++) {
		const int x = (int) (idx / sX);
		const int rem = (int) (idx % sX);
		const int y = rem / Z, z = rem % Z;
		const float lookup_x = (float) x + w[3 * idx];
		const float lookup_y = (float) y + w[3 * idx + 1];
		const float lookup_z = (float) z + w[3 * idx + 2];
		const int bx = (int) std::floor(lookup_x);
		const int by = (int) std::floor(lookup_y);
		const int bz = (int) std::floor(lookup_z);
		const float rx = lookup_x - (float) bx, ry = lookup_y - (float) by, rz = lookup_z - (float) bz;
		const float ix = 1.0f - rx, iy = 1.0f - ry, iz = 1.0f - rz;
		const bool x0 = bx >= 0 && bx < X, x1 = bx + 1 >= 0 && bx + 1 < X;
		const bool y0 = by >= 0 && by < Y, y1 = by + 1 >= 0 && by + 1 < Y;
		const bool z0 = bz >= 0 && bz < Z, z1 = bz + 1 >= 0 && bz + 1 < Z;
		const long b = (long) bx * sX + (long) by * sY + bz;
		for (int c = 0; c < C; c++) {
			const float v000 = (x0 && y0 && z0) ? f[(b) * C + c] : oob;
			const float v001 = (x0 && y0 && z1) ? f[(b + 1) * C + c] : oob;
			const float v010 = (x0 && y1 && z0) ? f[(b + sY) * C + c] : oob;
			const float v011 = (x0 && y1 && z1) ? f[(b + sY + 1) * C + c] : oob;
			const float v100 = (x1 && y0 && z0) ? f[(b + sX) * C + c] : oob;
			const float v101 = (x1 && y0 && z1) ? f[(b + sX + 1) * C + c] : oob;
			const float v110 = (x1 && y1 && z0) ? f[(b + sX + sY) * C + c] : oob;
			const float v111 = (x1 && y1 && z1) ? f[(b + sX + sY + 1) * C + c] : oob;
			// z, then y, then x -- field_warping.tpp:126-134
			const float i00 = v000 * iz + v001 * rz;
			const float i01 = v010 * iz + v011 * rz;
			const float i10 = v100 * iz + v101 * rz;
			const float i11 = v110 * iz + v111 * rz;
			const float i0 = i00 * iy + i01 * ry;
			const float i1 = i10 * iy + i11 * ry;
			out[idx * C + c] = i0 * ix + i1 * rx;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Central-difference gradient, one-sided at the borders. Reference: cpp/src/math/gradients.tpp:248-283
// (2D, .x = d/dcol, .y = d/drow), :438-495 (3D, component c = d/d(axis c)).
// ------------------------------------------------------------------------------------------------
inline float central_or_one_sided(const float* line, long stride, int i, int n) {
	if (i == 0) return line[stride] - line[0];
	if (i == n - 1) return line[(long) (n - 1) * stride] - line[(long) (n - 2) * stride];
	return 0.5f * (line[(long) (i + 1) * stride] - line[(long) (i - 1) * stride]);
}

void gradient2d(const float* f, int H, int W, float* out) {
#pragma omp parallel for
	for (int r = 0; r < H; r++) {
		for (int c = 0; c < W; c++) {
			const long idx = (long) r * W + c;
			out[2 * idx + 0] = central_or_one_sided(f + (long) r * W, 1, c, W);
			out[2 * idx + 1] = central_or_one_sided(f + c, W, r, H);
		}
	}
}

void gradient3d(const float* f, int X, int Y, int Z, float* out) {
	const long sX = (long) Y * Z, sY = Z;
#pragma omp parallel for
	for (int x = 0; x < X; x++) {
		for (int y = 0; y < Y; y++) {
			for (int z = 0; z < Z; z++) {
				const long idx = x * sX + y * sY + z;
				out[3 * idx + 0] = central_or_one_sided(f + y * sY + z, sX, x, X);
				out[3 * idx + 1] = central_or_one_sided(f + x * sX + z, sY, y, Y);
				out[3 * idx + 2] = central_or_one_sided(f + x * sX + y * sY, 1, z, Z);
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Laplacian of a vector field with replicated border. Reference: cpp/src/math/gradients.tpp:28-35
// (operators), :62-101 (2D: rows term assigned, columns term added), :106-172 (3D: axis 0 assigned,
// axis 1 added, axis 2 added). Interior term is (next - 2*cur) + prev; border term is
// nonborder - border.
// ------------------------------------------------------------------------------------------------
inline float laplace_axis_term(const float* line, long stride, int i, int n) {
	if (i == 0) return line[stride] - line[0];
	if (i == n - 1) return line[(long) (n - 2) * stride] - line[(long) (n - 1) * stride];
	return (line[(long) (i + 1) * stride] - 2 * line[(long) i * stride]) + line[(long) (i - 1) * stride];
}

void laplacian2d(const float* v, int C, int H, int W, float* out) {
#pragma omp parallel for
	for (int r = 0; r < H; r++) {
		for (int c = 0; c < W; c++) {
			for (int k = 0; k < C; k++) {
				float acc = laplace_axis_term(v + (long) c * C + k, (long) W * C, r, H);
				acc += laplace_axis_term(v + (long) r * W * C + k, C, c, W);
				out[((long) r * W + c) * C + k] = acc;
			}
		}
	}
}

void laplacian3d(const float* v, int C, int X, int Y, int Z, float* out) {
	const long sX = (long) Y * Z * C, sY = (long) Z * C, sZ = C;
#pragma omp parallel for
	for (int x = 0; x < X; x++) {
		for (int y = 0; y < Y; y++) {
			for (int z = 0; z < Z; z++) {
				for (int k = 0; k < C; k++) {
					float acc = laplace_axis_term(v + y * sY + z * sZ + k, sX, x, X);
					acc += laplace_axis_term(v + x * sX + z * sZ + k, sY, y, Y);
					acc += laplace_axis_term(v + x * sX + y * sY + k, sZ, z, Z);
					out[x * sX + y * sY + z * sZ + k] = acc;
				}
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Separable convolution, "same" size, zero padded, kernel flipped; accumulation from 0.0f over taps
// i-r .. i+r in ascending order. Reference: cpp/src/math/convolution.cpp:50-67 (helper), :147-219
// (2D: rows/y pass then columns/x pass), :221-332 (3D: axis 0, 1, 2), :23-47,:69-145 (preserve zeros:
// output is the zero vector where the pass-input vector is exactly zero).
// ------------------------------------------------------------------------------------------------
void convolve_axis(const float* in, float* out, int C, long outer_count, const long* outer_offsets, int n,
		long stride, const float* kernel, int K, bool preserve_zeros) {
	const int r = K / 2;
#pragma omp parallel for
	for (long o = 0; o < outer_count; o++) {
		const float* line_in = in + outer_offsets[o];
		float* line_out = out + outer_offsets[o];
		for (int i = 0; i < n; i++) {
			if (preserve_zeros) {
				bool all_zero = true;
				for (int c = 0; c < C; c++) all_zero = all_zero && (line_in[(long) i * stride + c] == 0.0f);
				if (all_zero) {
					for (int c = 0; c < C; c++) line_out[(long) i * stride + c] = 0.0f;
					continue;
				}
			}
			for (int c = 0; c < C; c++) {
				float acc = 0.0f;
				for (int j = 0; j < K; j++) {
					const int src = i - r + j;
					const float value = (src >= 0 && src < n) ? line_in[(long) src * stride + c] : 0.0f;
					acc += value * kernel[K - 1 - j];
				}
				line_out[(long) i * stride + c] = acc;
			}
		}
	}
}

void convolve2d(float* v, int C, int H, int W, const float* kernel, int K, bool preserve_zeros) {
	Field tmp((size_t) H * W * C);
	std::vector<long> offs;
	// pass 1: along rows index (axis 0), one line per column
	offs.resize(W);
	for (int c = 0; c < W; c++) offs[c] = (long) c * C;
	convolve_axis(v, tmp.data(), C, W, offs.data(), H, (long) W * C, kernel, K, preserve_zeros);
	// pass 2: along columns index (axis 1), one line per row
	offs.resize(H);
	for (int r = 0; r < H; r++) offs[r] = (long) r * W * C;
	convolve_axis(tmp.data(), v, C, H, offs.data(), W, C, kernel, K, preserve_zeros);
}

void convolve3d(float* v, int C, int X, int Y, int Z, const float* kernel, int K) {
	const long sX = (long) Y * Z * C, sY = (long) Z * C, sZ = C;
	Field tmp((size_t) X * Y * Z * C);
	std::vector<long> offs;
	offs.resize((size_t) Y * Z);
	for (int y = 0; y < Y; y++) for (int z = 0; z < Z; z++) offs[(size_t) y * Z + z] = y * sY + z * sZ;
	convolve_axis(v, tmp.data(), C, (long) Y * Z, offs.data(), X, sX, kernel, K, false);
	offs.resize((size_t) X * Z);
	for (int x = 0; x < X; x++) for (int z = 0; z < Z; z++) offs[(size_t) x * Z + z] = x * sX + z * sZ;
	convolve_axis(tmp.data(), v, C, (long) X * Z, offs.data(), Y, sY, kernel, K, false);
	offs.resize((size_t) X * Y);
	for (int x = 0; x < X; x++) for (int y = 0; y < Y; y++) offs[(size_t) x * Y + y] = x * sX + y * sY;
	convolve_axis(v, tmp.data(), C, (long) X * Y, offs.data(), Z, sZ, kernel, K, false);
	std::memcpy(v, tmp.data(), tmp.size() * sizeof(float));
}

// ------------------------------------------------------------------------------------------------
// max ||v|| : sqrt(max_i sum_c v_c^2), sum accumulated from 0 with c ascending.
// Reference: cpp/src/math/statistics.tpp:57-73 (2D), :76-100 (3D), vector_operations.hpp:50-55.
// ------------------------------------------------------------------------------------------------
float max_norm(const float* v, int C, long n) {
	float max_sq = 0.0f;
#pragma omp parallel for reduction(max:max_sq)
	for (long i = 0; i < n; i++) {
		float sq = 0.0f;
		for (int c = 0; c < C; c++) sq += v[i * C + c] * v[i * C + c];
		if (sq > max_sq) max_sq = sq;
	}
	return std::sqrt(max_sq);
}

// ------------------------------------------------------------------------------------------------
// Restrict x2. AVERAGE: cpp/src/math/resampling.tpp:358-382 (2D: (r,c)+(r,c+1)+(r+1,c)+(r+1,c+1), /4)
// and :385-417 (3D: first index fastest, /8). LINEAR 3D: :544-656 (4^3 tent on a replicate-padded
// field). LINEAR 2D: :422-541 (explicit border cases).
// ------------------------------------------------------------------------------------------------
void downsample2d_average(const float* f, int C, int H, int W, float* out) {
	const int h = H / 2, w = W / 2;
#pragma omp parallel for
	for (int r = 0; r < h; r++) {
		for (int c = 0; c < w; c++) {
			for (int k = 0; k < C; k++) {
				const float* p = f + ((long) (2 * r) * W + 2 * c) * C + k;
				out[((long) r * w + c) * C + k] = (p[0] + p[C] + p[(long) W * C] + p[(long) W * C + C]) / 4.0f;
			}
		}
	}
}

void downsample3d_average(const float* f, int C, int X, int Y, int Z, float* out) {
	const int dx = X / 2, dy = Y / 2, dz = Z / 2;
	const long sX = (long) Y * Z * C, sY = (long) Z * C, sZ = C;
#pragma omp parallel for
	for (int x = 0; x < dx; x++) {
		for (int y = 0; y < dy; y++) {
			for (int z = 0; z < dz; z++) {
				for (int k = 0; k < C; k++) {
					const float* p = f + (2 * x) * sX + (2 * y) * sY + (2 * z) * sZ + k;
					out[(((long) x * dy + y) * dz + z) * C + k] = (p[0] + p[sX] + p[sY] + p[sX + sY]
							+ p[sZ] + p[sX + sZ] + p[sY + sZ] + p[sX + sY + sZ]) / 8.0f;
				}
			}
		}
	}
}

inline int clampi(int v, int lo, int hi) {
	return v < lo ? lo : (v > hi ? hi : v);
}

void downsample3d_linear(const float* f, int C, int X, int Y, int Z, float* out) {
	const int dx = X / 2, dy = Y / 2, dz = Z / 2;
	const long sX = (long) Y * Z * C, sY = (long) Z * C, sZ = C;
	const float c0 = 0.052734375f * 4.0f, c1 = 0.017578125f * 4.0f, c2 = 0.005859375f * 4.0f, c3 = 0.001953125f
			* 4.0f;
	// tap offsets in the reference's summation order (resampling.tpp:566-649), relative to (2t) per axis
	static const int g0[8][3] = { { 0, 0, 0 }, { 1, 0, 0 }, { 0, 1, 0 }, { 1, 1, 0 }, { 0, 0, 1 }, { 1, 0, 1 }, { 0, 1,
			1 }, { 1, 1, 1 } };
	static const int g1[24][3] = { { -1, 0, 0 }, { 0, -1, 0 }, { 0, 0, -1 }, { 2, 0, 0 }, { 1, -1, 0 }, { 1, 0, -1 },
			{ -1, 1, 0 }, { 0, 2, 0 }, { 0, 1, -1 }, { 2, 1, 0 }, { 1, 2, 0 }, { 1, 1, -1 }, { -1, 0, 1 }, { 0, -1, 1 },
			{ 0, 0, 2 }, { 2, 0, 1 }, { 1, -1, 1 }, { 1, 0, 2 }, { -1, 1, 1 }, { 0, 2, 1 }, { 0, 1, 2 }, { 2, 1, 1 }, {
					1, 2, 1 }, { 1, 1, 2 } };
	static const int g2[24][3] = { { -1, -1, 0 }, { 0, -1, -1 }, { -1, 0, -1 }, { 2, -1, 0 }, { 1, -1, -1 }, { 2, 0, -1 },
			{ -1, 2, 0 }, { 0, 2, -1 }, { -1, 1, -1 }, { 2, 2, 0 }, { 1, 2, -1 }, { 2, 1, -1 }, { -1, -1, 1 }, { 0, -1,
					2 }, { -1, 0, 2 }, { 2, -1, 1 }, { 1, -1, 2 }, { 2, 0, 2 }, { -1, 2, 1 }, { 0, 2, 2 }, { -1, 1, 2 },
			{ 2, 2, 1 }, { 1, 2, 2 }, { 2, 1, 2 } };
	static const int g3[8][3] = { { -1, -1, -1 }, { 2, -1, -1 }, { -1, 2, -1 }, { 2, 2, -1 }, { -1, -1, 2 },
			{ 2, -1, 2 }, { -1, 2, 2 }, { 2, 2, 2 } };
#pragma omp parallel for
	for (int x = 0; x < dx; x++) {
		for (int y = 0; y < dy; y++) {
			for (int z = 0; z < dz; z++) {
				for (int k = 0; k < C; k++) {
					auto at = [&](const int* o) {
						const int xi = clampi(2 * x + o[0], 0, X - 1);
						const int yi = clampi(2 * y + o[1], 0, Y - 1);
						const int zi = clampi(2 * z + o[2], 0, Z - 1);
						return f[xi * sX + yi * sY + zi * sZ + k];
					};
					float s0 = at(g0[0]);
					for (int t = 1; t < 8; t++) s0 = s0 + at(g0[t]);
					float s1 = at(g1[0]);
					for (int t = 1; t < 24; t++) s1 = s1 + at(g1[t]);
					float s2 = at(g2[0]);
					for (int t = 1; t < 24; t++) s2 = s2 + at(g2[t]);
					float s3 = at(g3[0]);
					for (int t = 1; t < 8; t++) s3 = s3 + at(g3[t]);
					out[(((long) x * dy + y) * dz + z) * C + k] = (((c0 * s0 + c1 * s1) + c2 * s2) + c3 * s3) * 0.25f;
				}
			}
		}
	}
}

// 2D linear restrict, reference cpp/src/math/resampling.tpp:422-541. The reference writes the four
// corners, the two border rows, the two border columns and the interior as separate formulas whose
// tap ORDER differs; they are transcribed as offset tables (dr, dc) relative to the even source
// index (2*tr, 2*tc), mirrored for the far borders.
struct Tap {
	int dr, dc;
};

float sum_taps(const float* f, int C, int k, int W, int r0, int c0, int sr, int sc, const Tap* taps, int count) {
	// sr/sc = +1 for near border / interior, -1 for mirrored (far) border; r0/c0 = anchor
	float s = f[((long) (r0 + sr * taps[0].dr) * W + (c0 + sc * taps[0].dc)) * C + k];
	for (int t = 1; t < count; t++) {
		s = s + f[((long) (r0 + sr * taps[t].dr) * W + (c0 + sc * taps[t].dc)) * C + k];
	}
	return s;
}

void downsample2d_linear(const float* f, int C, int H, int W, float* out) {
	const int h = H / 2, w = W / 2;
	const float coeff0 = 0.140625f, coeff1 = 0.046875f, coeff2 = 0.015625f;
	const int lr = H - 1, lc = W - 1, dlr = h - 1, dlc = w - 1;
	// corner (anchored at the corner element, mirrored by sr/sc): resampling.tpp:438-461
	static const Tap corner0[4] = { { 0, 0 }, { 1, 0 }, { 0, 1 }, { 1, 1 } };
	static const Tap corner1[8] = { { 0, 0 }, { 0, 0 }, { 0, 1 }, { 1, 0 }, { 0, 2 }, { 1, 2 }, { 2, 1 }, { 2, 0 } };
	static const Tap corner2[4] = { { 0, 0 }, { 0, 2 }, { 2, 0 }, { 2, 2 } };
	// border rows (anchor row = border row, anchor col = source_col; mirrored in r): :467-491
	static const Tap brow0[4] = { { 0, 0 }, { 0, 1 }, { 1, 0 }, { 1, 1 } };
	static const Tap brow1[8] = { { 0, -1 }, { 0, 0 }, { 0, 1 }, { 0, 2 }, { 1, -1 }, { 2, 0 }, { 2, 1 }, { 1, 2 } };
	static const Tap brow2[4] = { { 0, -1 }, { 0, 2 }, { 2, -1 }, { 2, 2 } };
	// border columns (anchor col = border col, anchor row = source_row; mirrored in c): :497-519
	static const Tap bcol0[4] = { { 0, 0 }, { 1, 0 }, { 0, 1 }, { 1, 1 } };
	static const Tap bcol1[8] = { { -1, 0 }, { 0, 0 }, { 1, 0 }, { 2, 0 }, { -1, 1 }, { 0, 2 }, { 1, 2 }, { 2, 1 } };
	static const Tap bcol2[4] = { { -1, 0 }, { 2, 0 }, { -1, 2 }, { 2, 2 } };
	// interior: :523-537
	static const Tap in0[4] = { { 0, 0 }, { 0, 1 }, { 1, 0 }, { 1, 1 } };
	static const Tap in1[8] = { { -1, 0 }, { 0, -1 }, { -1, 1 }, { 0, 2 }, { 2, 0 }, { 1, -1 }, { 2, 1 }, { 1, 2 } };
	static const Tap in2[4] = { { -1, -1 }, { -1, 2 }, { 2, -1 }, { 2, 2 } };

	auto combine = [&](float s0, float s1, float s2) {
		return (coeff0 * s0 + coeff1 * s1) + coeff2 * s2;
	};
	for (int k = 0; k < C; k++) {
		struct {
			int tr, tc, r0, c0, sr, sc;
		} corners[4] = { { 0, 0, 0, 0, 1, 1 }, { 0, dlc, 0, lc, 1, -1 }, { dlr, 0, lr, 0, -1, 1 }, { dlr, dlc, lr, lc,
				-1, -1 } };
		for (auto& cn : corners) {
			out[((long) cn.tr * w + cn.tc) * C + k] = combine(
					sum_taps(f, C, k, W, cn.r0, cn.c0, cn.sr, cn.sc, corner0, 4),
					sum_taps(f, C, k, W, cn.r0, cn.c0, cn.sr, cn.sc, corner1, 8),
					sum_taps(f, C, k, W, cn.r0, cn.c0, cn.sr, cn.sc, corner2, 4));
		}
		for (int tc = 1; tc < w - 1; tc++) {
			const int sc = tc * 2;
			out[((long) 0 * w + tc) * C + k] = combine(sum_taps(f, C, k, W, 0, sc, 1, 1, brow0, 4),
					sum_taps(f, C, k, W, 0, sc, 1, 1, brow1, 8), sum_taps(f, C, k, W, 0, sc, 1, 1, brow2, 4));
			out[((long) dlr * w + tc) * C + k] = combine(sum_taps(f, C, k, W, lr, sc, -1, 1, brow0, 4),
					sum_taps(f, C, k, W, lr, sc, -1, 1, brow1, 8), sum_taps(f, C, k, W, lr, sc, -1, 1, brow2, 4));
		}
		for (int tr = 1; tr < h - 1; tr++) {
			const int sr = tr * 2;
			out[((long) tr * w + 0) * C + k] = combine(sum_taps(f, C, k, W, sr, 0, 1, 1, bcol0, 4),
					sum_taps(f, C, k, W, sr, 0, 1, 1, bcol1, 8), sum_taps(f, C, k, W, sr, 0, 1, 1, bcol2, 4));
			out[((long) tr * w + dlc) * C + k] = combine(sum_taps(f, C, k, W, sr, lc, 1, -1, bcol0, 4),
					sum_taps(f, C, k, W, sr, lc, 1, -1, bcol1, 8), sum_taps(f, C, k, W, sr, lc, 1, -1, bcol2, 4));
		}
		for (int tr = 1; tr < h - 1; tr++) {
			for (int tc = 1; tc < w - 1; tc++) {
				out[((long) tr * w + tc) * C + k] = combine(sum_taps(f, C, k, W, 2 * tr, 2 * tc, 1, 1, in0, 4),
						sum_taps(f, C, k, W, 2 * tr, 2 * tc, 1, 1, in1, 8),
						sum_taps(f, C, k, W, 2 * tr, 2 * tc, 1, 1, in2, 4));
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Prolong x2. NEAREST: resampling.tpp:68-82 (2D), :103-126 (3D). LINEAR 2D: :130-216 (border rows and
// columns 0.75/0.25 along the border, corners copied, interior 4-weight formula). LINEAR 3D: :218-322
// (six faces are 2D-upsampled source faces, later faces overwrite earlier ones on shared edges;
// interior separable 0.75/0.25 along axis 0, then 1, then 2).
// ------------------------------------------------------------------------------------------------
void upsample2d_nearest(const float* f, int C, int H, int W, float* out) {
#pragma omp parallel for
	for (int r = 0; r < 2 * H; r++)
		for (int c = 0; c < 2 * W; c++)
			for (int k = 0; k < C; k++) out[((long) r * 2 * W + c) * C + k] = f[((long) (r / 2) * W + c / 2) * C + k];
}

void upsample3d_nearest(const float* f, int C, int X, int Y, int Z, float* out) {
#pragma omp parallel for
	for (int x = 0; x < 2 * X; x++)
		for (int y = 0; y < 2 * Y; y++)
			for (int z = 0; z < 2 * Z; z++)
				for (int k = 0; k < C; k++)
					out[(((long) x * 2 * Y + y) * 2 * Z + z) * C + k] = f[(((long) (x / 2) * Y + y / 2) * Z + z / 2) * C
							+ k];
}

// strided 2D linear upsample: source element (r,c) at f[r*srs + c*scs], target at out[r*trs + c*tcs]
void upsample2d_linear_strided(const float* f, long srs, long scs, int H, int W, float* out, long trs, long tcs) {
	const int UH = 2 * H, UW = 2 * W;
	auto S = [&](int r, int c) {return f[r * srs + c * scs];};
	auto T = [&](int r, int c) -> float& {return out[r * trs + c * tcs];};
	// first and last rows incl. corners
	T(0, 0) = S(0, 0);
	T(UH - 1, 0) = S(H - 1, 0);
	float p0 = S(0, 0), p1 = S(H - 1, 0);
	for (int sc = 1, dc = 1; sc < W; sc++, dc += 2) {
		const float c0 = S(0, sc), c1 = S(H - 1, sc);
		T(0, dc) = 0.75f * p0 + 0.25f * c0;
		T(0, dc + 1) = 0.25f * p0 + 0.75f * c0;
		T(UH - 1, dc) = 0.75f * p1 + 0.25f * c1;
		T(UH - 1, dc + 1) = 0.25f * p1 + 0.75f * c1;
		p0 = c0;
		p1 = c1;
	}
	T(0, UW - 1) = p0;
	T(UH - 1, UW - 1) = p1;
	// first and last columns excl. corners
	p0 = S(0, 0);
	p1 = S(0, W - 1);
	for (int sr = 1, dr = 1; sr < H; sr++, dr += 2) {
		const float c0 = S(sr, 0), c1 = S(sr, W - 1);
		T(dr, 0) = 0.75f * p0 + 0.25f * c0;
		T(dr + 1, 0) = 0.25f * p0 + 0.75f * c0;
		T(dr, UW - 1) = 0.75f * p1 + 0.25f * c1;
		T(dr + 1, UW - 1) = 0.25f * p1 + 0.75f * c1;
		p0 = c0;
		p1 = c1;
	}
	// midsection
	for (int sc = 0; sc < W - 1; sc++) {
		const int dc = 1 + 2 * sc;
		for (int sr = 1, dr = 1; sr < H; sr++, dr += 2) {
			const float v00 = S(sr - 1, sc), v01 = S(sr - 1, sc + 1), v10 = S(sr, sc), v11 = S(sr, sc + 1);
			T(dr, dc) = ((0.5625f * v00 + 0.1875f * v01) + 0.1875f * v10) + 0.0625f * v11;
			T(dr, dc + 1) = ((0.1875f * v00 + 0.5625f * v01) + 0.0625f * v10) + 0.1875f * v11;
			T(dr + 1, dc) = ((0.1875f * v00 + 0.0625f * v01) + 0.5625f * v10) + 0.1875f * v11;
			T(dr + 1, dc + 1) = ((0.0625f * v00 + 0.1875f * v01) + 0.1875f * v10) + 0.5625f * v11;
		}
	}
}

void upsample2d_linear(const float* f, int C, int H, int W, float* out) {
	for (int k = 0; k < C; k++)
		upsample2d_linear_strided(f + k, (long) W * C, C, H, W, out + k, (long) 2 * W * C, C);
}

void upsample3d_linear(const float* f, int C, int X, int Y, int Z, float* out) {
	const long sX = (long) Y * Z * C, sY = (long) Z * C, sZ = C;
	const int UX = 2 * X, UY = 2 * Y, UZ = 2 * Z;
	const long tX = (long) UY * UZ * C, tY = (long) UZ * C, tZ = C;
	for (int k = 0; k < C; k++) {
		// Faces in the reference's order (resampling.tpp:255-275). The reference maps a (1,Y,Z) slice to a
		// column-major Y x Z matrix: matrix rows <-> first remaining tensor index, columns <-> second.
		upsample2d_linear_strided(f + k, sY, sZ, Y, Z, out + k, tY, tZ);                                    // near x
		upsample2d_linear_strided(f + (X - 1) * sX + k, sY, sZ, Y, Z, out + (UX - 1) * tX + k, tY, tZ);     // far x
		upsample2d_linear_strided(f + k, sX, sZ, X, Z, out + k, tX, tZ);                                    // near y
		upsample2d_linear_strided(f + (Y - 1) * sY + k, sX, sZ, X, Z, out + (UY - 1) * tY + k, tX, tZ);     // far y
		upsample2d_linear_strided(f + k, sX, sY, X, Y, out + k, tX, tY);                                    // near z
		upsample2d_linear_strided(f + (Z - 1) * sZ + k, sX, sY, X, Y, out + (UZ - 1) * tZ + k, tX, tY);     // far z
	}
#pragma omp parallel for
	for (int x = 0; x < X - 1; x++) {
		for (int y = 0; y < Y - 1; y++) {
			for (int z = 0; z < Z - 1; z++) {
				for (int k = 0; k < C; k++) {
					const float* p = f + x * sX + y * sY + z * sZ + k;
					const float v000 = p[0], v100 = p[sX], v010 = p[sY], v110 = p[sX + sY];
					const float v001 = p[sZ], v101 = p[sX + sZ], v011 = p[sY + sZ], v111 = p[sX + sY + sZ];
					const float xv000 = 0.75f * v000 + 0.25f * v100, xv100 = 0.25f * v000 + 0.75f * v100;
					const float xv010 = 0.75f * v010 + 0.25f * v110, xv110 = 0.25f * v010 + 0.75f * v110;
					const float xv001 = 0.75f * v001 + 0.25f * v101, xv101 = 0.25f * v001 + 0.75f * v101;
					const float xv011 = 0.75f * v011 + 0.25f * v111, xv111 = 0.25f * v011 + 0.75f * v111;
					const float yv000 = 0.75f * xv000 + 0.25f * xv010, yv010 = 0.25f * xv000 + 0.75f * xv010;
					const float yv100 = 0.75f * xv100 + 0.25f * xv110, yv110 = 0.25f * xv100 + 0.75f * xv110;
					const float yv001 = 0.75f * xv001 + 0.25f * xv011, yv011 = 0.25f * xv001 + 0.75f * xv011;
					const float yv101 = 0.75f * xv101 + 0.25f * xv111, yv111 = 0.25f * xv101 + 0.75f * xv111;
					float* q = out + (2 * x + 1) * tX + (2 * y + 1) * tY + (2 * z + 1) * tZ + k;
					q[0] = 0.75f * yv000 + 0.25f * yv001;
					q[tX] = 0.75f * yv100 + 0.25f * yv101;
					q[tY] = 0.75f * yv010 + 0.25f * yv011;
					q[tX + tY] = 0.75f * yv110 + 0.25f * yv111;
					q[tZ] = 0.25f * yv000 + 0.75f * yv001;
					q[tX + tZ] = 0.25f * yv100 + 0.75f * yv101;
					q[tY + tZ] = 0.25f * yv010 + 0.75f * yv011;
					q[tX + tY + tZ] = 0.25f * yv110 + 0.75f * yv111;
				}
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Hierarchical optimizer. Reference: cpp/src/nonrigid_optimization/hierarchical/optimizer.tpp:83-131
// (optimize), :134-171 (level loop + termination), :174-212 (iteration); pyramid.tpp:51-74.
// ------------------------------------------------------------------------------------------------
struct Dims {
	int d[3];
	int nd;
	long voxels() const {
		long n = 1;
		for (int i = 0; i < nd; i++) n *= d[i];
		return n;
	}
};

int pyramid_level_count(const orc_hier_params* p, const Dims& dims) {
	if (!is_power_of_two(p->maximum_chunk_size)) return -1;
	const int power = (int) std::log2((double) p->maximum_chunk_size);
	double min_log = 1e30;
	for (int i = 0; i < dims.nd; i++) min_log = std::min(min_log, std::log2((double) dims.d[i]));
	const int max_level_count = (int) min_log + 1;
	if (!(max_level_count > power)) return -2;
	return power + 1;
}

void downsample(const Field& in, int C, const Dims& dims, bool linear, Field& out, Dims& out_dims) {
	out_dims = dims;
	for (int i = 0; i < dims.nd; i++) out_dims.d[i] = dims.d[i] / 2;
	out.resize((size_t) out_dims.voxels() * C);
	if (dims.nd == 2) {
		if (linear) downsample2d_linear(in.data(), C, dims.d[0], dims.d[1], out.data());
		else downsample2d_average(in.data(), C, dims.d[0], dims.d[1], out.data());
	} else {
		if (linear) downsample3d_linear(in.data(), C, dims.d[0], dims.d[1], dims.d[2], out.data());
		else downsample3d_average(in.data(), C, dims.d[0], dims.d[1], dims.d[2], out.data());
	}
}

struct Pyramid {
	std::vector<Field> levels; // coarsest first
	std::vector<Dims> dims;
	Pyramid(const float* field, int C, const Dims& d, int level_count, bool linear) {
		levels.emplace_back(field, field + d.voxels() * C);
		dims.push_back(d);
		for (int i = 1; i < level_count; i++) {
			Field next;
			Dims nd;
			downsample(levels.back(), C, dims.back(), linear, next, nd);
			levels.push_back(std::move(next));
			dims.push_back(nd);
		}
		std::reverse(levels.begin(), levels.end());
		std::reverse(dims.begin(), dims.end());
	}
};

struct IterationBuffers {
	Field resampled_live, resampled_gradient, laplacian, gradient;
};

float hier_iteration(const orc_hier_params* p, bool tikhonov, bool use_kernel, const Dims& d, const float* canonical,
		const float* live, const float* live_gradient, Field& warp, IterationBuffers& b) {
	const int D = d.nd;
	const long n = d.voxels();
	b.resampled_live.resize(n);
	b.resampled_gradient.resize((size_t) n * D);
	if (D == 2) {
		warp2d_generic(live, 1, warp.data(), d.d[0], d.d[1], 1.0f, b.resampled_live.data());
		warp2d_generic(live_gradient, 2, warp.data(), d.d[0], d.d[1], 0.0f, b.resampled_gradient.data());
	} else {
		warp3d_generic(live, 1, warp.data(), d.d[0], d.d[1], d.d[2], 1.0f, b.resampled_live.data());
		warp3d_generic(live_gradient, 3, warp.data(), d.d[0], d.d[1], d.d[2], 0.0f, b.resampled_gradient.data());
	}
	if (tikhonov) {
		b.laplacian.resize((size_t) n * D);
		if (D == 2) laplacian2d(b.gradient.data(), 2, d.d[0], d.d[1], b.laplacian.data());
		else laplacian3d(b.gradient.data(), 3, d.d[0], d.d[1], d.d[2], b.laplacian.data());
	}
	const float amp = p->data_term_amplifier, strength = p->tikhonov_strength, rate = p->rate;
	float* g = b.gradient.data();
#pragma omp parallel for
	for (long i = 0; i < n; i++) {
		const float diff = b.resampled_live[i] - canonical[i];
		for (int c = 0; c < D; c++) {
			const float data_gradient = b.resampled_gradient[i * D + c] * diff;
			if (tikhonov) g[i * D + c] = data_gradient * amp - b.laplacian[i * D + c] * strength;
			else g[i * D + c] = data_gradient * amp;
		}
	}
	if (use_kernel) {
		if (D == 2) convolve2d(g, 2, d.d[0], d.d[1], p->kernel, p->kernel_size, false);
		else convolve3d(g, 3, d.d[0], d.d[1], d.d[2], p->kernel, p->kernel_size);
	}
	float* w = warp.data();
#pragma omp parallel for
	for (long i = 0; i < n * D; i++) w[i] = w[i] - g[i] * rate;
	return max_norm(g, D, n);
}

int hier_optimize(const orc_hier_params* p, const float* canonical, const float* live, const Dims& dims,
		float* warp_out, int* iteration_counts, float* max_update_lengths, orc_iteration_dump* dump) {
	const int D = dims.nd;
	const int level_count = pyramid_level_count(p, dims);
	if (level_count < 0) return level_count;
	const bool linear = p->resampling_strategy == 1;
	for (int i = 0; i < D; i++) {
		if (dims.d[i] % (1 << (level_count - 1)) != 0) return -3;
		if (!linear && D == 2 && !is_power_of_two(dims.d[i])) return -4;
	}
	const bool tikhonov = p->tikhonov_term_enabled && p->tikhonov_strength > 0.0f;
	const bool use_kernel = p->gradient_kernel_enabled && p->kernel_size > 0 && p->kernel != nullptr;

	Field live_gradient((size_t) dims.voxels() * D);
	if (D == 2) gradient2d(live, dims.d[0], dims.d[1], live_gradient.data());
	else gradient3d(live, dims.d[0], dims.d[1], dims.d[2], live_gradient.data());

	Pyramid canonical_pyramid(canonical, 1, dims, level_count, linear);
	Pyramid live_pyramid(live, 1, dims, level_count, linear);
	Pyramid gradient_pyramid(live_gradient.data(), D, dims, level_count, linear);

	Field warp;
	IterationBuffers buffers;
	if (dump) dump->count = 0;
	for (int level = 0; level < level_count; level++) {
		const Dims& d = canonical_pyramid.dims[level];
		const long n = d.voxels();
		if (level == 0) warp.assign((size_t) n * D, 0.0f);
		buffers.gradient.assign((size_t) n * D, 0.0f);
		float max_update = FLT_MAX;
		int iteration = 0;
		while (!(max_update < p->maximum_warp_update_threshold || iteration >= p->maximum_iteration_count)) {
			max_update = hier_iteration(p, tikhonov, use_kernel, d, canonical_pyramid.levels[level].data(),
					live_pyramid.levels[level].data(), gradient_pyramid.levels[level].data(), warp, buffers);
			if (dump && dump->level == level && dump->buffer && iteration < dump->max_iterations) {
				std::memcpy(dump->buffer + (size_t) iteration * n * D, warp.data(), (size_t) n * D * sizeof(float));
				dump->count = iteration + 1;
			}
			iteration++;
		}
		if (iteration_counts) iteration_counts[level] = iteration;
		if (max_update_lengths) max_update_lengths[level] = max_update;
		if (level != level_count - 1) {
			Field up((size_t) n * D * (D == 2 ? 4 : 8));
			if (D == 2) {
				if (linear) upsample2d_linear(warp.data(), 2, d.d[0], d.d[1], up.data());
				else upsample2d_nearest(warp.data(), 2, d.d[0], d.d[1], up.data());
			} else {
				if (linear) upsample3d_linear(warp.data(), 3, d.d[0], d.d[1], d.d[2], up.data());
				else upsample3d_nearest(warp.data(), 3, d.d[0], d.d[1], d.d[2], up.data());
			}
			warp.swap(up);
		}
	}
	std::memcpy(warp_out, warp.data(), warp.size() * sizeof(float));
	return level_count;
}

} // namespace

extern "C" {

void orc_warp2d(const float* field, const float* warp, int H, int W, float* out) {
	warp2d_generic(field, 1, warp, H, W, 1.0f, out);
}
void orc_warp2d_replacement(const float* field, int C, const float* warp, int H, int W, float replacement,
		float* out) {
	warp2d_generic(field, C, warp, H, W, replacement, out);
}
void orc_gradient2d(const float* field, int H, int W, float* out) {
	gradient2d(field, H, W, out);
}
void orc_laplacian2d(const float* vfield, int C, int H, int W, float* out) {
	laplacian2d(vfield, C, H, W, out);
}
void orc_convolve2d(float* vfield, int C, int H, int W, const float* kernel, int K, int preserve_zeros) {
	convolve2d(vfield, C, H, W, kernel, K, preserve_zeros != 0);
}
int orc_downsample2d(const float* field, int C, int H, int W, int linear, float* out) {
	if (linear) {
		if (H % 2 || W % 2 || H <= 2 || W <= 2) return -1;
		downsample2d_linear(field, C, H, W, out);
	} else {
		if (!is_power_of_two(H) || !is_power_of_two(W)) return -1;
		downsample2d_average(field, C, H, W, out);
	}
	return 0;
}
int orc_upsample2d(const float* field, int C, int H, int W, int linear, float* out) {
	if (linear) upsample2d_linear(field, C, H, W, out);
	else upsample2d_nearest(field, C, H, W, out);
	return 0;
}
float orc_max_norm2d(const float* vfield, int C, long n) {
	return max_norm(vfield, C, n);
}

void orc_warp3d(const float* field, const float* warp, int X, int Y, int Z, float* out) {
	warp3d_generic(field, 1, warp, X, Y, Z, 1.0f, out);
}
void orc_warp3d_replacement(const float* field, int C, const float* warp, int X, int Y, int Z, float replacement,
		float* out) {
	warp3d_generic(field, C, warp, X, Y, Z, replacement, out);
}
void orc_gradient3d(const float* field, int X, int Y, int Z, float* out) {
	gradient3d(field, X, Y, Z, out);
}
void orc_laplacian3d(const float* vfield, int C, int X, int Y, int Z, float* out) {
	laplacian3d(vfield, C, X, Y, Z, out);
}
void orc_convolve3d(float* vfield, int C, int X, int Y, int Z, const float* kernel, int K) {
	convolve3d(vfield, C, X, Y, Z, kernel, K);
}
int orc_downsample3d(const float* field, int C, int X, int Y, int Z, int linear, float* out) {
	if (linear) {
		if (X % 2 || Y % 2 || Z % 2 || X <= 2 || Y <= 2 || Z <= 2) return -1;
		downsample3d_linear(field, C, X, Y, Z, out);
	} else {
		downsample3d_average(field, C, X, Y, Z, out);
	}
	return 0;
}
int orc_upsample3d(const float* field, int C, int X, int Y, int Z, int linear, float* out) {
	if (linear) upsample3d_linear(field, C, X, Y, Z, out);
	else upsample3d_nearest(field, C, X, Y, Z, out);
	return 0;
}

int orc_hier_optimize2d(const orc_hier_params* p, const float* canonical, const float* live, int H, int W,
		float* warp_out, int* iteration_counts, float* max_update_lengths, orc_iteration_dump* dump) {
	Dims d { { H, W, 1 }, 2 };
	return hier_optimize(p, canonical, live, d, warp_out, iteration_counts, max_update_lengths, dump);
}
int orc_hier_optimize3d(const orc_hier_params* p, const float* canonical, const float* live, int X, int Y, int Z,
		float* warp_out, int* iteration_counts, float* max_update_lengths, orc_iteration_dump* dump) {
	Dims d { { X, Y, Z }, 3 };
	return hier_optimize(p, canonical, live, d, warp_out, iteration_counts, max_update_lengths, dump);
}

double orc_hier_time_iterations3d(const orc_hier_params* p, const float* canonical, const float* live, int X, int Y,
		int Z, int iterations) {
	Dims d { { X, Y, Z }, 3 };
	const long n = d.voxels();
	Field live_gradient((size_t) n * 3);
	gradient3d(live, X, Y, Z, live_gradient.data());
	Field warp((size_t) n * 3, 0.0f);
	IterationBuffers buffers;
	buffers.gradient.assign((size_t) n * 3, 0.0f);
	const bool tikhonov = p->tikhonov_term_enabled && p->tikhonov_strength > 0.0f;
	const bool use_kernel = p->gradient_kernel_enabled && p->kernel_size > 0 && p->kernel != nullptr;
	// one untimed warm-up iteration (page faults of the temporaries)
	hier_iteration(p, tikhonov, use_kernel, d, canonical, live, live_gradient.data(), warp, buffers);
	const double t0 = omp_get_wtime();
	for (int i = 0; i < iterations; i++)
		hier_iteration(p, tikhonov, use_kernel, d, canonical, live, live_gradient.data(), warp, buffers);
	return omp_get_wtime() - t0;
}

void orc_set_num_threads(int n) {
	if (n > 0) omp_set_num_threads(n);
}

int orc_num_threads(void) {
	return omp_get_max_threads();
}

} // extern "C"
