/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the reference Gaussian-splatting hot path.
 *
 * Plain C11, float32 arithmetic written in the reference's order of operations. Every function cites
 * the reference file:line it restates (paths relative to /root/reference). Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the product
 * (photo-slam_b200/) never does.
 *
 * PINNING STATUS: the reference ships no tests / golden vectors for this path. This port
 * is pinned against outputs of the reference's own kernels (oracle/_ref, compiled unmodified from
 * /root/reference and run on a B200): tests/golden/*.npz were dumped by tests/golden/make_golden.py on
 * the GPU box and tests/test_oracle_golden.py checks this file against them on CPU.
 *
 * Known, documented deviation from the GPU reference: nvcc/ptxas decide FMA contraction per expression;
 * gcc is run with -ffp-contract=off and only the contractions that are stable and visible in the
 * reference's SASS are restated with explicit fmaf() (point transforms, det, eigenvalue discriminant,
 * the 3-term dot products of the covariance chain). libm expf() also differs from the device expf() in
 * the last ulp. Integer outputs (radii, tiles, keys, n_contrib) therefore agree with the GPU reference
 * except for a handful of threshold cases per million; tests state the allowed count explicitly.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BLOCK_X 16
#define BLOCK_Y 16

/* ---- SH constants: cuda_rasterizer/auxiliary.h:22-39 ---- */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

/* a0*b0 + a1*b1 + a2*b2 as the device evaluates it: fma(a2,b2, fma(a0,b0, a1*b1)) (reference SASS). */
static inline float dot3f(float a0, float b0, float a1, float b1, float a2, float b2)
{
	return fmaf(a2, b2, fmaf(a0, b0, a1 * b1));
}

/* cuda_rasterizer/auxiliary.h:41-44 — evaluated in double, narrowed to float */
static inline float ndc2Pix(float v, int S) { return (float)((((double)v + 1.0) * (double)S - 1.0) * 0.5); }

/* cuda_rasterizer/auxiliary.h:58-77 (column-major 4x4: m[4*c+r]) */
static inline void transformPoint4x3(const float* p, const float* m, float* o)
{
	o[0] = dot3f(m[0], p[0], m[4], p[1], m[8], p[2]) + m[12];
	o[1] = dot3f(m[1], p[0], m[5], p[1], m[9], p[2]) + m[13];
	o[2] = dot3f(m[2], p[0], m[6], p[1], m[10], p[2]) + m[14];
}
static inline void transformPoint4x4(const float* p, const float* m, float* o)
{
	o[0] = dot3f(m[0], p[0], m[4], p[1], m[8], p[2]) + m[12];
	o[1] = dot3f(m[1], p[0], m[5], p[1], m[9], p[2]) + m[13];
	o[2] = dot3f(m[2], p[0], m[6], p[1], m[10], p[2]) + m[14];
	o[3] = dot3f(m[3], p[0], m[7], p[1], m[11], p[2]) + m[15];
}

/* cuda_rasterizer/auxiliary.h:46-56 */
static inline void getRect(float px, float py, int max_radius, int gx, int gy, int* rmin, int* rmax)
{
	int v;
	v = (int)((px - (float)max_radius) / (float)BLOCK_X); if (v < 0) v = 0; if (v > gx) v = gx; rmin[0] = v;
	v = (int)((py - (float)max_radius) / (float)BLOCK_Y); if (v < 0) v = 0; if (v > gy) v = gy; rmin[1] = v;
	v = (int)((px + (float)max_radius + (float)BLOCK_X - 1.0f) / (float)BLOCK_X); if (v < 0) v = 0; if (v > gx) v = gx; rmax[0] = v;
	v = (int)((py + (float)max_radius + (float)BLOCK_Y - 1.0f) / (float)BLOCK_Y); if (v < 0) v = 0; if (v > gy) v = gy; rmax[1] = v;
}

/* 3x3 column-major helpers with glm semantics: M[c][r] stored at m[3*c+r];
 * (A*B)[c][r] = A[0][r]*B[c][0] + A[1][r]*B[c][1] + A[2][r]*B[c][2]. */
static inline void mat3_mul(const float* a, const float* b, float* r)
{
	for (int c = 0; c < 3; c++)
		for (int rr = 0; rr < 3; rr++)
			r[3 * c + rr] = dot3f(a[0 + rr], b[3 * c + 0], a[3 + rr], b[3 * c + 1], a[6 + rr], b[3 * c + 2]);
}
static inline void mat3_transpose(const float* a, float* r)
{
	for (int c = 0; c < 3; c++) for (int rr = 0; rr < 3; rr++) r[3 * c + rr] = a[3 * rr + c];
}

/* forward.cu:118-152 computeCov3D. rot is (r,x,y,z), NOT normalised here. */
static void computeCov3D(const float* scale, float mod, const float* rot, float* cov3D)
{
	float S[9] = {0}; S[0] = mod * scale[0]; S[4] = mod * scale[1]; S[8] = mod * scale[2];
	float r = rot[0], x = rot[1], y = rot[2], z = rot[3];
	float R[9] = {
		1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
		2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
		2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)};
	float M[9], Mt[9], Sigma[9];
	mat3_mul(S, R, M);
	mat3_transpose(M, Mt);
	mat3_mul(Mt, M, Sigma);
	cov3D[0] = Sigma[0]; cov3D[1] = Sigma[1]; cov3D[2] = Sigma[2];
	cov3D[3] = Sigma[4]; cov3D[4] = Sigma[5]; cov3D[5] = Sigma[8];
}

/* Shared by forward.cu:74-113 (computeCov2D) and backward.cu:155-197: builds T = W*J and cov2D. */
static void cov2d_common(const float* mean, float focal_x, float focal_y, float tan_fovx, float tan_fovy,
                         const float* cov3D, const float* vm, float* t_out, float* T, float* Vrk, float* cov,
                         float* x_grad_mul, float* y_grad_mul)
{
	float t[3];
	transformPoint4x3(mean, vm, t);
	const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
	const float txtz = t[0] / t[2], tytz = t[1] / t[2];
	t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
	t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
	*x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
	*y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
	float J[9] = {focal_x / t[2], 0.f, -(focal_x * t[0]) / (t[2] * t[2]),
	              0.f, focal_y / t[2], -(focal_y * t[1]) / (t[2] * t[2]),
	              0.f, 0.f, 0.f};
	float W[9] = {vm[0], vm[4], vm[8], vm[1], vm[5], vm[9], vm[2], vm[6], vm[10]};
	mat3_mul(W, J, T);
	float V[9] = {cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]};
	memcpy(Vrk, V, sizeof(V));
	float Tt[9], Vt[9], A[9];
	mat3_transpose(T, Tt);
	mat3_transpose(V, Vt);
	mat3_mul(Tt, Vt, A);
	mat3_mul(A, T, cov);
	t_out[0] = t[0]; t_out[1] = t[1]; t_out[2] = t[2];
}

/* forward.cu:20-71 computeColorFromSH. sh row: max_coeffs x 3 floats. */
static void computeColorFromSH(int deg, const float* pos, const float* campos, const float* sh, float* rgb, uint8_t* clamped)
{
	float dir[3] = {pos[0] - campos[0], pos[1] - campos[1], pos[2] - campos[2]};
	float len = sqrtf(dot3f(dir[0], dir[0], dir[1], dir[1], dir[2], dir[2]));
	dir[0] /= len; dir[1] /= len; dir[2] /= len;
	float x = dir[0], y = dir[1], z = dir[2];
	for (int ch = 0; ch < 3; ch++) {
#define SH(k) sh[3 * (k) + ch]
		float result = SH_C0 * SH(0);
		if (deg > 0) {
			result = result - SH_C1 * y * SH(1) + SH_C1 * z * SH(2) - SH_C1 * x * SH(3);
			if (deg > 1) {
				float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
				result = result + SH_C2[0] * xy * SH(4) + SH_C2[1] * yz * SH(5) +
				         SH_C2[2] * (2.0f * zz - xx - yy) * SH(6) + SH_C2[3] * xz * SH(7) + SH_C2[4] * (xx - yy) * SH(8);
				if (deg > 2) {
					result = result + SH_C3[0] * y * (3.0f * xx - yy) * SH(9) + SH_C3[1] * xy * z * SH(10) +
					         SH_C3[2] * y * (4.0f * zz - xx - yy) * SH(11) +
					         SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SH(12) +
					         SH_C3[4] * x * (4.0f * zz - xx - yy) * SH(13) + SH_C3[5] * z * (xx - yy) * SH(14) +
					         SH_C3[6] * x * (xx - 3.0f * yy) * SH(15);
				}
			}
		}
#undef SH
		result += 0.5f;
		clamped[ch] = (result < 0);
		rgb[ch] = result < 0.f ? 0.f : result;
	}
}

/*
 * forward.cu:155-256 preprocessCUDA (+ auxiliary.h:139-164 in_frustum).
 * Outputs for culled Gaussians are left untouched except radii = tiles_touched = 0;
 * callers zero-initialise the arrays so comparisons are well defined.
 */
void orc_preprocess(int P, int D, int M, const float* means3D, const float* scales, float scale_modifier,
                    const float* rotations, const float* opacities, const float* shs, const float* cov3D_precomp,
                    const float* colors_precomp, const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                    int W, int H, float tan_fovx, float tan_fovy,
                    int* radii, float* means2D, float* depths, float* cov3Ds, float* rgb, float* conic_opacity,
                    uint32_t* tiles_touched, uint8_t* clamped)
{
	const float focal_y = H / (2.0f * tan_fovy), focal_x = W / (2.0f * tan_fovx); /* rasterizer_impl.cu:221-222 */
	const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
#pragma omp parallel for schedule(static)
	for (int idx = 0; idx < P; idx++) {
		radii[idx] = 0; tiles_touched[idx] = 0;
		const float* p_orig = means3D + 3 * idx;
		float p_view[3];
		transformPoint4x3(p_orig, viewmatrix, p_view);
		if (p_view[2] <= 0.2f) continue;
		float p_hom[4];
		transformPoint4x4(p_orig, projmatrix, p_hom);
		float p_w = 1.0f / (p_hom[3] + 0.0000001f);
		float p_proj[3] = {p_hom[0] * p_w, p_hom[1] * p_w, p_hom[2] * p_w};
		const float* cov3D;
		if (cov3D_precomp) cov3D = cov3D_precomp + 6 * idx;
		else { computeCov3D(scales + 3 * idx, scale_modifier, rotations + 4 * idx, cov3Ds + 6 * idx); cov3D = cov3Ds + 6 * idx; }
		float t[3], T[9], Vrk[9], cov2[9], xg, yg;
		cov2d_common(p_orig, focal_x, focal_y, tan_fovx, tan_fovy, cov3D, viewmatrix, t, T, Vrk, cov2, &xg, &yg);
		float cx = cov2[0] + 0.3f, cy = cov2[1], cz = cov2[4] + 0.3f; /* cov[0][0], cov[0][1], cov[1][1] */
		float det = fmaf(cx, cz, -(cy * cy));
		if (det == 0.0f) continue;
		float det_inv = 1.f / det;
		float conic[3] = {cz * det_inv, -cy * det_inv, cx * det_inv};
		float mid = 0.5f * (cx + cz);
		float disc = fmaxf(0.1f, fmaf(mid, mid, -det));
		float lambda1 = mid + sqrtf(disc), lambda2 = mid - sqrtf(disc);
		float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
		float px = ndc2Pix(p_proj[0], W), py = ndc2Pix(p_proj[1], H);
		int rmin[2], rmax[2];
		getRect(px, py, (int)my_radius, gx, gy, rmin, rmax);
		if ((rmax[0] - rmin[0]) * (rmax[1] - rmin[1]) == 0) continue;
		if (!colors_precomp)
			computeColorFromSH(D, p_orig, cam_pos, shs + (size_t)idx * M * 3, rgb + 3 * idx, clamped + 3 * idx);
		depths[idx] = p_view[2];
		radii[idx] = (int)my_radius;
		means2D[2 * idx] = px; means2D[2 * idx + 1] = py;
		conic_opacity[4 * idx] = conic[0]; conic_opacity[4 * idx + 1] = conic[1];
		conic_opacity[4 * idx + 2] = conic[2]; conic_opacity[4 * idx + 3] = opacities[idx];
		tiles_touched[idx] = (uint32_t)((rmax[1] - rmin[1]) * (rmax[0] - rmin[0]));
	}
}

/* rasterizer_impl.cu:141-153 + :54-66 (checkFrustum) */
void orc_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present)
{
	(void)projmatrix;
	for (int i = 0; i < P; i++) {
		float pv[3];
		transformPoint4x3(means3D + 3 * i, viewmatrix, pv);
		present[i] = pv[2] > 0.2f;
	}
}

typedef struct { uint64_t key; uint32_t val; uint32_t seq; } kv_t;
static int kv_cmp(const void* a, const void* b)
{
	const kv_t* x = (const kv_t*)a; const kv_t* y = (const kv_t*)b;
	if (x->key < y->key) return -1;
	if (x->key > y->key) return 1;
	return (x->seq < y->seq) ? -1 : (x->seq > y->seq);
}

/*
 * Binning: rasterizer_impl.cu:274-318 (inclusive scan, duplicateWithKeys :70-111, stable radix sort on
 * (tile<<32 | depth bits), identifyTileRanges :116-138). Returns num_rendered; if keys == NULL only counts.
 * point_offsets[P] = inclusive scan. keys/values sized >= num_rendered. ranges = 2*T uint32, zero for untouched tiles.
 */
int orc_bin(int P, int W, int H, const int* radii, const float* means2D, const float* depths,
            const uint32_t* tiles_touched, uint32_t* point_offsets, uint64_t* keys, uint32_t* values, uint32_t* ranges)
{
	const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
	uint32_t run = 0;
	for (int i = 0; i < P; i++) { run += tiles_touched[i]; point_offsets[i] = run; }
	int N = (int)run;
	if (!keys) return N;
	kv_t* kv = (kv_t*)malloc(sizeof(kv_t) * (size_t)(N > 0 ? N : 1));
	for (int idx = 0; idx < P; idx++) {
		if (radii[idx] <= 0) continue;
		uint32_t off = idx == 0 ? 0 : point_offsets[idx - 1];
		int rmin[2], rmax[2];
		getRect(means2D[2 * idx], means2D[2 * idx + 1], radii[idx], gx, gy, rmin, rmax);
		uint32_t dbits; memcpy(&dbits, &depths[idx], 4);
		for (int y = rmin[1]; y < rmax[1]; y++)
			for (int x = rmin[0]; x < rmax[0]; x++) {
				uint64_t key = (uint64_t)(y * gx + x);
				key <<= 32; key |= dbits;
				kv[off].key = key; kv[off].val = (uint32_t)idx; kv[off].seq = off; off++;
			}
	}
	qsort(kv, (size_t)N, sizeof(kv_t), kv_cmp);
	memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)gx * gy);
	for (int i = 0; i < N; i++) {
		keys[i] = kv[i].key; values[i] = kv[i].val;
		uint32_t cur = (uint32_t)(kv[i].key >> 32);
		if (i == 0) ranges[2 * cur] = 0;
		else {
			uint32_t prev = (uint32_t)(kv[i - 1].key >> 32);
			if (cur != prev) { ranges[2 * prev + 1] = (uint32_t)i; ranges[2 * cur] = (uint32_t)i; }
		}
		if (i == N - 1) ranges[2 * cur + 1] = (uint32_t)N;
	}
	free(kv);
	return N;
}

/* forward.cu:261-374 renderCUDA: per pixel, front-to-back over the tile's list. colors: [P,3]. out_color CHW. */
void orc_render_forward(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* means2D,
                        const float* colors, const float* conic_opacity, const float* bg,
                        float* out_color, float* final_T, uint32_t* n_contrib)
{
	const int gx = (W + BLOCK_X - 1) / BLOCK_X;
#pragma omp parallel for schedule(dynamic, 4)
	for (int py = 0; py < H; py++)
		for (int px = 0; px < W; px++) {
			int tile = (py / BLOCK_Y) * gx + (px / BLOCK_X);
			uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
			float pixfx = (float)px, pixfy = (float)py;
			float T = 1.0f, C[3] = {0, 0, 0};
			uint32_t contributor = 0, last_contributor = 0;
			for (uint32_t k = r0; k < r1; k++) {
				contributor++;
				uint32_t g = point_list[k];
				float dx = means2D[2 * g] - pixfx, dy = means2D[2 * g + 1] - pixfy;
				const float* co = conic_opacity + 4 * g;
				float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
				if (power > 0.0f) continue;
				float alpha = fminf(0.99f, co[3] * expf(power));
				if (alpha < 1.0f / 255.0f) continue;
				float test_T = T * (1 - alpha);
				if (test_T < 0.0001f) break; /* done = true */
				for (int ch = 0; ch < 3; ch++) C[ch] += colors[3 * g + ch] * alpha * T;
				T = test_T;
				last_contributor = contributor;
			}
			size_t pid = (size_t)py * W + px;
			final_T[pid] = T;
			n_contrib[pid] = last_contributor;
			for (int ch = 0; ch < 3; ch++) out_color[(size_t)ch * H * W + pid] = C[ch] + T * bg[ch];
		}
}

/*
 * backward.cu:399-557 renderCUDA: per pixel, back-to-front. Per-pixel arithmetic in float as the reference;
 * the cross-pixel sum (atomicAdd in the reference, order undefined) is accumulated in double and rounded once.
 * dL_dmean2D [P,3] (x,y written), dL_dconic [P,4] (x,y,w written), dL_dopacity [P], dL_dcolors [P,3].
 */
void orc_render_backward(int P, int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* bg,
                         const float* means2D, const float* conic_opacity, const float* colors,
                         const float* final_Ts, const uint32_t* n_contrib, const float* dL_dpixels,
                         float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolors)
{
	const int gx = (W + BLOCK_X - 1) / BLOCK_X;
	double* acc = (double*)calloc((size_t)P * 9, sizeof(double));
	const float ddelx_dx = (float)(0.5 * W), ddely_dy = (float)(0.5 * H);
	for (int py = 0; py < H; py++)
		for (int px = 0; px < W; px++) {
			int tile = (py / BLOCK_Y) * gx + (px / BLOCK_X);
			uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
			size_t pid = (size_t)py * W + px;
			const float T_final = final_Ts[pid];
			float T = T_final;
			const uint32_t last_contributor = n_contrib[pid];
			float accum_rec[3] = {0, 0, 0}, dL_dpixel[3], last_alpha = 0, last_color[3] = {0, 0, 0};
			for (int ch = 0; ch < 3; ch++) dL_dpixel[ch] = dL_dpixels[(size_t)ch * H * W + pid];
			float pixfx = (float)px, pixfy = (float)py;
			uint32_t toDo = r1 - r0;
			uint32_t contributor = toDo;
			for (uint32_t j = 0; j < toDo; j++) {
				contributor--;
				if (contributor >= last_contributor) continue;
				uint32_t g = point_list[r1 - j - 1];
				float dx = means2D[2 * g] - pixfx, dy = means2D[2 * g + 1] - pixfy;
				const float* co = conic_opacity + 4 * g;
				float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
				if (power > 0.0f) continue;
				const float G = expf(power);
				const float alpha = fminf(0.99f, co[3] * G);
				if (alpha < 1.0f / 255.0f) continue;
				T = T / (1.f - alpha);
				const float dchannel_dcolor = alpha * T;
				float dL_dalpha = 0.0f;
				for (int ch = 0; ch < 3; ch++) {
					const float c = colors[3 * g + ch];
					accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
					last_color[ch] = c;
					const float dL_dchannel = dL_dpixel[ch];
					dL_dalpha += (c - accum_rec[ch]) * dL_dchannel;
					acc[(size_t)g * 9 + 6 + ch] += (double)(dchannel_dcolor * dL_dchannel);
				}
				dL_dalpha *= T;
				last_alpha = alpha;
				float bg_dot_dpixel = 0;
				for (int i = 0; i < 3; i++) bg_dot_dpixel += bg[i] * dL_dpixel[i];
				dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;
				const float dL_dG = co[3] * dL_dalpha;
				const float gdx = G * dx, gdy = G * dy;
				const float dG_ddelx = -gdx * co[0] - gdy * co[1];
				const float dG_ddely = -gdy * co[2] - gdx * co[1];
				acc[(size_t)g * 9 + 0] += (double)(dL_dG * dG_ddelx * ddelx_dx);
				acc[(size_t)g * 9 + 1] += (double)(dL_dG * dG_ddely * ddely_dy);
				acc[(size_t)g * 9 + 2] += (double)(-0.5f * gdx * dx * dL_dG);
				acc[(size_t)g * 9 + 3] += (double)(-0.5f * gdx * dy * dL_dG);
				acc[(size_t)g * 9 + 4] += (double)(-0.5f * gdy * dy * dL_dG);
				acc[(size_t)g * 9 + 5] += (double)(G * dL_dalpha);
			}
		}
	for (int g = 0; g < P; g++) {
		const double* a = acc + (size_t)g * 9;
		dL_dmean2D[3 * g] += (float)a[0]; dL_dmean2D[3 * g + 1] += (float)a[1];
		dL_dconic[4 * g] += (float)a[2]; dL_dconic[4 * g + 1] += (float)a[3]; dL_dconic[4 * g + 3] += (float)a[4];
		dL_dopacity[g] += (float)a[5];
		for (int ch = 0; ch < 3; ch++) dL_dcolors[3 * g + ch] += (float)a[6 + ch];
	}
	free(acc);
}

/* auxiliary.h:107-117 dnormvdv (float3) */
static void dnormvdv3(const float* v, const float* dv, float* o)
{
	float sum2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
	float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
	o[0] = ((+sum2 - v[0] * v[0]) * dv[0] - v[1] * v[0] * dv[1] - v[2] * v[0] * dv[2]) * invsum32;
	o[1] = (-v[0] * v[1] * dv[0] + (sum2 - v[1] * v[1]) * dv[1] - v[2] * v[1] * dv[2]) * invsum32;
	o[2] = (-v[0] * v[2] * dv[0] - v[1] * v[2] * dv[1] + (sum2 - v[2] * v[2]) * dv[2]) * invsum32;
}

/* backward.cu:20-139 computeColorFromSH (backward). dL_dsh row: M x 3. dL_dmean += view-direction term. */
static void shBackward(int deg, const float* pos, const float* campos, const float* sh, const uint8_t* clamped,
                       const float* dL_dcolor, float* dL_dmean, float* dL_dsh)
{
	float dir_orig[3] = {pos[0] - campos[0], pos[1] - campos[1], pos[2] - campos[2]};
	float len = sqrtf(dot3f(dir_orig[0], dir_orig[0], dir_orig[1], dir_orig[1], dir_orig[2], dir_orig[2]));
	float x = dir_orig[0] / len, y = dir_orig[1] / len, z = dir_orig[2] / len;
	float dL_dRGB[3];
	for (int ch = 0; ch < 3; ch++) dL_dRGB[ch] = dL_dcolor[ch] * (clamped[ch] ? 0.f : 1.f);
	float dRGBdx[3] = {0, 0, 0}, dRGBdy[3] = {0, 0, 0}, dRGBdz[3] = {0, 0, 0};
#define SH(k) sh[3 * (k) + ch]
#define DSH(k, w) dL_dsh[3 * (k) + ch] = (w) * dL_dRGB[ch]
	for (int ch = 0; ch < 3; ch++) {
		DSH(0, SH_C0);
		if (deg > 0) {
			DSH(1, -SH_C1 * y); DSH(2, SH_C1 * z); DSH(3, -SH_C1 * x);
			dRGBdx[ch] = -SH_C1 * SH(3); dRGBdy[ch] = -SH_C1 * SH(1); dRGBdz[ch] = SH_C1 * SH(2);
			if (deg > 1) {
				float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
				DSH(4, SH_C2[0] * xy); DSH(5, SH_C2[1] * yz); DSH(6, SH_C2[2] * (2.f * zz - xx - yy));
				DSH(7, SH_C2[3] * xz); DSH(8, SH_C2[4] * (xx - yy));
				dRGBdx[ch] += SH_C2[0] * y * SH(4) + SH_C2[2] * 2.f * -x * SH(6) + SH_C2[3] * z * SH(7) + SH_C2[4] * 2.f * x * SH(8);
				dRGBdy[ch] += SH_C2[0] * x * SH(4) + SH_C2[1] * z * SH(5) + SH_C2[2] * 2.f * -y * SH(6) + SH_C2[4] * 2.f * -y * SH(8);
				dRGBdz[ch] += SH_C2[1] * y * SH(5) + SH_C2[2] * 2.f * 2.f * z * SH(6) + SH_C2[3] * x * SH(7);
				if (deg > 2) {
					DSH(9, SH_C3[0] * y * (3.f * xx - yy)); DSH(10, SH_C3[1] * xy * z);
					DSH(11, SH_C3[2] * y * (4.f * zz - xx - yy)); DSH(12, SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
					DSH(13, SH_C3[4] * x * (4.f * zz - xx - yy)); DSH(14, SH_C3[5] * z * (xx - yy));
					DSH(15, SH_C3[6] * x * (xx - 3.f * yy));
					dRGBdx[ch] += (SH_C3[0] * SH(9) * 3.f * 2.f * xy + SH_C3[1] * SH(10) * yz + SH_C3[2] * SH(11) * -2.f * xy +
					               SH_C3[3] * SH(12) * -3.f * 2.f * xz + SH_C3[4] * SH(13) * (-3.f * xx + 4.f * zz - yy) +
					               SH_C3[5] * SH(14) * 2.f * xz + SH_C3[6] * SH(15) * 3.f * (xx - yy));
					dRGBdy[ch] += (SH_C3[0] * SH(9) * 3.f * (xx - yy) + SH_C3[1] * SH(10) * xz +
					               SH_C3[2] * SH(11) * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * SH(12) * -3.f * 2.f * yz +
					               SH_C3[4] * SH(13) * -2.f * xy + SH_C3[5] * SH(14) * -2.f * yz + SH_C3[6] * SH(15) * -3.f * 2.f * xy);
					dRGBdz[ch] += (SH_C3[1] * SH(10) * xy + SH_C3[2] * SH(11) * 4.f * 2.f * yz +
					               SH_C3[3] * SH(12) * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * SH(13) * 4.f * 2.f * xz +
					               SH_C3[5] * SH(14) * (xx - yy));
				}
			}
		}
	}
#undef SH
#undef DSH
	float dL_ddir[3] = {
		dRGBdx[0] * dL_dRGB[0] + dRGBdx[1] * dL_dRGB[1] + dRGBdx[2] * dL_dRGB[2],
		dRGBdy[0] * dL_dRGB[0] + dRGBdy[1] * dL_dRGB[1] + dRGBdy[2] * dL_dRGB[2],
		dRGBdz[0] * dL_dRGB[0] + dRGBdz[1] * dL_dRGB[1] + dRGBdz[2] * dL_dRGB[2]};
	float dm[3];
	dnormvdv3(dir_orig, dL_ddir, dm);
	dL_dmean[0] += dm[0]; dL_dmean[1] += dm[1]; dL_dmean[2] += dm[2];
}

/* backward.cu:278-341 computeCov3D (backward) */
static void cov3DBackward(const float* scale, float mod, const float* rot, const float* dL_dcov3D, float* dL_dscale, float* dL_drot)
{
	float r = rot[0], x = rot[1], y = rot[2], z = rot[3];
	float R[9] = {
		1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
		2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
		2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)};
	float s[3] = {mod * scale[0], mod * scale[1], mod * scale[2]};
	float S[9] = {0}; S[0] = s[0]; S[4] = s[1]; S[8] = s[2];
	float M[9]; mat3_mul(S, R, M);
	float dL_dSigma[9] = {
		dL_dcov3D[0], 0.5f * dL_dcov3D[1], 0.5f * dL_dcov3D[2],
		0.5f * dL_dcov3D[1], dL_dcov3D[3], 0.5f * dL_dcov3D[4],
		0.5f * dL_dcov3D[2], 0.5f * dL_dcov3D[4], dL_dcov3D[5]};
	float M2[9]; for (int i = 0; i < 9; i++) M2[i] = M[i] * 2.0f; /* 2.0f * M (scalar * mat: column * scalar) */
	float dL_dM[9]; mat3_mul(M2, dL_dSigma, dL_dM);
	float Rt[9], dL_dMt[9];
	mat3_transpose(R, Rt); mat3_transpose(dL_dM, dL_dMt);
	for (int c = 0; c < 3; c++)
		dL_dscale[c] = Rt[3 * c] * dL_dMt[3 * c] + Rt[3 * c + 1] * dL_dMt[3 * c + 1] + Rt[3 * c + 2] * dL_dMt[3 * c + 2];
	for (int c = 0; c < 3; c++) for (int rr = 0; rr < 3; rr++) dL_dMt[3 * c + rr] *= s[c];
#define D(c, rr) dL_dMt[3 * (c) + (rr)]
	dL_drot[0] = 2 * z * (D(0, 1) - D(1, 0)) + 2 * y * (D(2, 0) - D(0, 2)) + 2 * x * (D(1, 2) - D(2, 1));
	dL_drot[1] = 2 * y * (D(1, 0) + D(0, 1)) + 2 * z * (D(2, 0) + D(0, 2)) + 2 * r * (D(1, 2) - D(2, 1)) - 4 * x * (D(2, 2) + D(1, 1));
	dL_drot[2] = 2 * x * (D(1, 0) + D(0, 1)) + 2 * r * (D(2, 0) - D(0, 2)) + 2 * z * (D(1, 2) + D(2, 1)) - 4 * y * (D(2, 2) + D(0, 0));
	dL_drot[3] = 2 * r * (D(0, 1) - D(1, 0)) + 2 * x * (D(2, 0) + D(0, 2)) + 2 * y * (D(1, 2) + D(2, 1)) - 4 * z * (D(1, 1) + D(0, 0));
#undef D
}

/*
 * BACKWARD::preprocess (backward.cu:559-621): computeCov2DCUDA (:144-274) then preprocessCUDA (:346-396).
 * All gradient outputs are accumulated into / assigned exactly as the reference; rows with radii <= 0 untouched.
 * dL_dconic [P,4] (x,y,w read), dL_dmean2D [P,3], dL_dcolor [P,3] (in), outputs: dL_dmeans [P,3], dL_dcov3D [P,6],
 * dL_dsh [P,M,3], dL_dscale [P,3], dL_drot [P,4].
 */
void orc_preprocess_backward(int P, int D, int M, const float* means3D, const int* radii, const float* shs,
                             const uint8_t* clamped, const float* scales, const float* rotations, float scale_modifier,
                             const float* cov3Ds, const float* viewmatrix, const float* projmatrix,
                             int W, int H, float tan_fovx, float tan_fovy, const float* campos,
                             const float* dL_dmean2D, const float* dL_dconic, float* dL_dmeans, const float* dL_dcolor,
                             float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot)
{
	const float h_y = H / (2.0f * tan_fovy), h_x = W / (2.0f * tan_fovx);
	const float* proj = projmatrix;
#pragma omp parallel for schedule(static)
	for (int idx = 0; idx < P; idx++) {
		if (!(radii[idx] > 0)) continue;
		const float* mean = means3D + 3 * idx;
		const float* cov3D = cov3Ds + 6 * idx;
		float dc[3] = {dL_dconic[4 * idx], dL_dconic[4 * idx + 1], dL_dconic[4 * idx + 3]};
		float t[3], T[9], Vrk[9], cov2D[9], x_grad_mul, y_grad_mul;
		cov2d_common(mean, h_x, h_y, tan_fovx, tan_fovy, cov3D, viewmatrix, t, T, Vrk, cov2D, &x_grad_mul, &y_grad_mul);
		float W3[9] = {viewmatrix[0], viewmatrix[4], viewmatrix[8], viewmatrix[1], viewmatrix[5], viewmatrix[9],
		               viewmatrix[2], viewmatrix[6], viewmatrix[10]};
		float a = cov2D[0] + 0.3f, b = cov2D[1], c = cov2D[4] + 0.3f;
		float denom = a * c - b * b;
		float dL_da = 0, dL_db = 0, dL_dc = 0;
		float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
#define Tm(c_, r_) T[3 * (c_) + (r_)]
#define Vm(c_, r_) Vrk[3 * (c_) + (r_)]
#define Wm(c_, r_) W3[3 * (c_) + (r_)]
		float* dcov = dL_dcov3D + 6 * idx;
		if (denom2inv != 0) {
			dL_da = denom2inv * (-c * c * dc[0] + 2 * b * c * dc[1] + (denom - a * c) * dc[2]);
			dL_dc = denom2inv * (-a * a * dc[2] + 2 * a * b * dc[1] + (denom - a * c) * dc[0]);
			dL_db = denom2inv * 2 * (b * c * dc[0] - (denom + 2 * b * b) * dc[1] + a * b * dc[2]);
			dcov[0] = (Tm(0, 0) * Tm(0, 0) * dL_da + Tm(0, 0) * Tm(1, 0) * dL_db + Tm(1, 0) * Tm(1, 0) * dL_dc);
			dcov[3] = (Tm(0, 1) * Tm(0, 1) * dL_da + Tm(0, 1) * Tm(1, 1) * dL_db + Tm(1, 1) * Tm(1, 1) * dL_dc);
			dcov[5] = (Tm(0, 2) * Tm(0, 2) * dL_da + Tm(0, 2) * Tm(1, 2) * dL_db + Tm(1, 2) * Tm(1, 2) * dL_dc);
			dcov[1] = 2 * Tm(0, 0) * Tm(0, 1) * dL_da + (Tm(0, 0) * Tm(1, 1) + Tm(0, 1) * Tm(1, 0)) * dL_db + 2 * Tm(1, 0) * Tm(1, 1) * dL_dc;
			dcov[2] = 2 * Tm(0, 0) * Tm(0, 2) * dL_da + (Tm(0, 0) * Tm(1, 2) + Tm(0, 2) * Tm(1, 0)) * dL_db + 2 * Tm(1, 0) * Tm(1, 2) * dL_dc;
			dcov[4] = 2 * Tm(0, 2) * Tm(0, 1) * dL_da + (Tm(0, 1) * Tm(1, 2) + Tm(0, 2) * Tm(1, 1)) * dL_db + 2 * Tm(1, 1) * Tm(1, 2) * dL_dc;
		} else {
			for (int i = 0; i < 6; i++) dcov[i] = 0;
		}
		float dL_dT00 = 2 * (Tm(0, 0) * Vm(0, 0) + Tm(0, 1) * Vm(0, 1) + Tm(0, 2) * Vm(0, 2)) * dL_da +
		                (Tm(1, 0) * Vm(0, 0) + Tm(1, 1) * Vm(0, 1) + Tm(1, 2) * Vm(0, 2)) * dL_db;
		float dL_dT01 = 2 * (Tm(0, 0) * Vm(1, 0) + Tm(0, 1) * Vm(1, 1) + Tm(0, 2) * Vm(1, 2)) * dL_da +
		                (Tm(1, 0) * Vm(1, 0) + Tm(1, 1) * Vm(1, 1) + Tm(1, 2) * Vm(1, 2)) * dL_db;
		float dL_dT02 = 2 * (Tm(0, 0) * Vm(2, 0) + Tm(0, 1) * Vm(2, 1) + Tm(0, 2) * Vm(2, 2)) * dL_da +
		                (Tm(1, 0) * Vm(2, 0) + Tm(1, 1) * Vm(2, 1) + Tm(1, 2) * Vm(2, 2)) * dL_db;
		float dL_dT10 = 2 * (Tm(1, 0) * Vm(0, 0) + Tm(1, 1) * Vm(0, 1) + Tm(1, 2) * Vm(0, 2)) * dL_dc +
		                (Tm(0, 0) * Vm(0, 0) + Tm(0, 1) * Vm(0, 1) + Tm(0, 2) * Vm(0, 2)) * dL_db;
		float dL_dT11 = 2 * (Tm(1, 0) * Vm(1, 0) + Tm(1, 1) * Vm(1, 1) + Tm(1, 2) * Vm(1, 2)) * dL_dc +
		                (Tm(0, 0) * Vm(1, 0) + Tm(0, 1) * Vm(1, 1) + Tm(0, 2) * Vm(1, 2)) * dL_db;
		float dL_dT12 = 2 * (Tm(1, 0) * Vm(2, 0) + Tm(1, 1) * Vm(2, 1) + Tm(1, 2) * Vm(2, 2)) * dL_dc +
		                (Tm(0, 0) * Vm(2, 0) + Tm(0, 1) * Vm(2, 1) + Tm(0, 2) * Vm(2, 2)) * dL_db;
		float dL_dJ00 = Wm(0, 0) * dL_dT00 + Wm(0, 1) * dL_dT01 + Wm(0, 2) * dL_dT02;
		float dL_dJ02 = Wm(2, 0) * dL_dT00 + Wm(2, 1) * dL_dT01 + Wm(2, 2) * dL_dT02;
		float dL_dJ11 = Wm(1, 0) * dL_dT10 + Wm(1, 1) * dL_dT11 + Wm(1, 2) * dL_dT12;
		float dL_dJ12 = Wm(2, 0) * dL_dT10 + Wm(2, 1) * dL_dT11 + Wm(2, 2) * dL_dT12;
#undef Tm
#undef Vm
#undef Wm
		float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
		float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
		float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
		float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t[0]) * tz3 * dL_dJ02 + (2 * h_y * t[1]) * tz3 * dL_dJ12;
		/* transformVec4x3Transpose, auxiliary.h:89-97 */
		float* dm = dL_dmeans + 3 * idx;
		dm[0] = viewmatrix[0] * dL_dtx + viewmatrix[1] * dL_dty + viewmatrix[2] * dL_dtz;
		dm[1] = viewmatrix[4] * dL_dtx + viewmatrix[5] * dL_dty + viewmatrix[6] * dL_dtz;
		dm[2] = viewmatrix[8] * dL_dtx + viewmatrix[9] * dL_dty + viewmatrix[10] * dL_dtz;

		/* backward.cu:366-395 */
		const float* m = mean;
		float m_hom[4];
		transformPoint4x4(m, proj, m_hom);
		float m_w = 1.0f / (m_hom[3] + 0.0000001f);
		float mul1 = (proj[0] * m[0] + proj[4] * m[1] + proj[8] * m[2] + proj[12]) * m_w * m_w;
		float mul2 = (proj[1] * m[0] + proj[5] * m[1] + proj[9] * m[2] + proj[13]) * m_w * m_w;
		const float gx2 = dL_dmean2D[3 * idx], gy2 = dL_dmean2D[3 * idx + 1];
		dm[0] += (proj[0] * m_w - proj[3] * mul1) * gx2 + (proj[1] * m_w - proj[3] * mul2) * gy2;
		dm[1] += (proj[4] * m_w - proj[7] * mul1) * gx2 + (proj[5] * m_w - proj[7] * mul2) * gy2;
		dm[2] += (proj[8] * m_w - proj[11] * mul1) * gx2 + (proj[9] * m_w - proj[11] * mul2) * gy2;
		if (shs)
			shBackward(D, m, campos, shs + (size_t)idx * M * 3, clamped + 3 * idx, dL_dcolor + 3 * idx, dm, dL_dsh + (size_t)idx * M * 3);
		if (scales)
			cov3DBackward(scales + 3 * idx, scale_modifier, rotations + 4 * idx, dcov, dL_dscale + 3 * idx, dL_drot + 4 * idx);
	}
}

/* ---------------- loss: include/loss_utils.h:28-124 (L1 + SSIM, 11x11 sigma 1.5, zero padding, groups=3) ---- */
static void gauss_window(float* w)
{
	float s = 0;
	for (int x = 0; x < 11; x++) { int t = x - 5; w[x] = expf(-(float)(t * t) / (2.0f * 1.5f * 1.5f)); s += w[x]; }
	for (int x = 0; x < 11; x++) w[x] /= s;
}
/* separable 11x11 zero-padded conv of one [H,W] plane (window = outer product of w with itself, as create_window) */
static void conv11(const float* in, float* out, int H, int W, const float* w, float* tmp)
{
	for (int y = 0; y < H; y++)
		for (int x = 0; x < W; x++) {
			double a = 0;
			for (int k = -5; k <= 5; k++) { int xx = x + k; if (xx >= 0 && xx < W) a += (double)w[k + 5] * in[(size_t)y * W + xx]; }
			tmp[(size_t)y * W + x] = (float)a;
		}
	for (int y = 0; y < H; y++)
		for (int x = 0; x < W; x++) {
			double a = 0;
			for (int k = -5; k <= 5; k++) { int yy = y + k; if (yy >= 0 && yy < H) a += (double)w[k + 5] * tmp[(size_t)yy * W + x]; }
			out[(size_t)y * W + x] = (float)a;
		}
}

/*
 * loss = (1-lambda)*mean|img-gt| + lambda*(1 - mean(ssim_map))   (gaussian_mapper.cpp:692-698)
 * Returns loss; writes l1, ssim (means) and dL/dimg [3,H,W] if dL_dimg != NULL (autograd of the same graph).
 */
float orc_loss(int H, int W, const float* img, const float* gt, float lambda_dssim, float* out_l1, float* out_ssim, float* dL_dimg)
{
	const size_t HW = (size_t)H * W;
	float w[11]; gauss_window(w);
	const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
	double l1 = 0, ss = 0;
	float* buf = (float*)malloc(sizeof(float) * HW * 12);
	float *mu1 = buf, *mu2 = buf + HW, *e11 = buf + 2 * HW, *e22 = buf + 3 * HW, *e12 = buf + 4 * HW, *tmp = buf + 5 * HW,
	      *prod = buf + 6 * HW, *g_m1 = buf + 7 * HW, *g_e11 = buf + 8 * HW, *g_e12 = buf + 9 * HW, *c1 = buf + 10 * HW, *c2 = buf + 11 * HW;
	const double inv_n = 1.0 / (3.0 * (double)HW);
	for (int ch = 0; ch < 3; ch++) {
		const float* x = img + ch * HW; const float* y = gt + ch * HW;
		conv11(x, mu1, H, W, w, tmp); conv11(y, mu2, H, W, w, tmp);
		for (size_t i = 0; i < HW; i++) prod[i] = x[i] * x[i];
		conv11(prod, e11, H, W, w, tmp);
		for (size_t i = 0; i < HW; i++) prod[i] = y[i] * y[i];
		conv11(prod, e22, H, W, w, tmp);
		for (size_t i = 0; i < HW; i++) prod[i] = x[i] * y[i];
		conv11(prod, e12, H, W, w, tmp);
		for (size_t i = 0; i < HW; i++) {
			double m1 = mu1[i], m2 = mu2[i];
			double s1 = e11[i] - m1 * m1, s2 = e22[i] - m2 * m2, s12 = e12[i] - m1 * m2;
			double A1 = 2 * m1 * m2 + C1, A2 = 2 * s12 + C2, B1 = m1 * m1 + m2 * m2 + C1, B2 = s1 + s2 + C2;
			double map = (A1 * A2) / (B1 * B2);
			ss += map;
			l1 += fabs((double)x[i] - (double)y[i]);
			if (dL_dimg) {
				double g = -(double)lambda_dssim * inv_n; /* dL/dmap */
				double dmap_dm1 = (2 * m2 * A2 + A1 * 2 * (-m2)) / (B1 * B2) - map * (2 * m1 / B1 + (-2 * m1) / B2);
				double dmap_de11 = -map / B2;
				double dmap_de12 = 2 * A1 / (B1 * B2);
				g_m1[i] = (float)(g * dmap_dm1); g_e11[i] = (float)(g * dmap_de11); g_e12[i] = (float)(g * dmap_de12);
			}
		}
		if (dL_dimg) {
			conv11(g_m1, c1, H, W, w, tmp);
			conv11(g_e11, c2, H, W, w, tmp);
			conv11(g_e12, prod, H, W, w, tmp);
			for (size_t i = 0; i < HW; i++) {
				float d = x[i] - y[i];
				float sgn = (d > 0) ? 1.f : ((d < 0) ? -1.f : 0.f);
				dL_dimg[ch * HW + i] = (float)((1.0 - lambda_dssim) * sgn * inv_n) + c1[i] + 2.f * x[i] * c2[i] + y[i] * prod[i];
			}
		}
	}
	free(buf);
	float fl1 = (float)(l1 * inv_n), fss = (float)(ss * inv_n);
	if (out_l1) *out_l1 = fl1;
	if (out_ssim) *out_ssim = fss;
	return (1.0f - lambda_dssim) * fl1 + lambda_dssim * (1.0f - fss);
}

/* ---------------- Adam: LibTorch 2.0.1 torch::optim::Adam::step semantics (call site
 * gaussian_mapper.cpp:769-772, setup gaussian_model.cpp:483-503): beta=(0.9,0.999), eps=1e-15, no weight decay/amsgrad.
 * step_t is the 1-based step count AFTER increment. */
void orc_adam(size_t n, float* p, const float* g, float* m, float* v, float lr, float beta1, float beta2, float eps, int step_t)
{
	const double bc1 = 1.0 - pow((double)beta1, step_t), bc2 = 1.0 - pow((double)beta2, step_t);
	const float step_size = (float)(lr / bc1);
	const float bc2_sqrt = (float)sqrt(bc2);
	for (size_t i = 0; i < n; i++) {
		m[i] = m[i] * beta1 + (1 - beta1) * g[i];            /* exp_avg.mul_(b1).add_(grad, 1-b1) */
		v[i] = v[i] * beta2 + (1 - beta2) * g[i] * g[i];      /* exp_avg_sq.mul_(b2).addcmul_(grad, grad, 1-b2) */
		float denom = sqrtf(v[i]) / bc2_sqrt + eps;         /* (exp_avg_sq.sqrt() / sqrt(bc2)).add_(eps) */
		p[i] = p[i] - step_size * (m[i] / denom);            /* p.addcdiv_(exp_avg, denom, -step_size) */
	}
}

/* ---------------- activations + their autograd (gaussian_model.cpp:48-71) ---------------- */
void orc_activations(int P, const float* opacity_raw, const float* scaling_raw, const float* rotation_raw,
                     float* opacity, float* scales, float* rotations)
{
	for (int i = 0; i < P; i++) {
		opacity[i] = 1.0f / (1.0f + expf(-opacity_raw[i]));
		for (int k = 0; k < 3; k++) scales[3 * i + k] = expf(scaling_raw[3 * i + k]);
		const float* q = rotation_raw + 4 * i;
		float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
		if (n < 1e-12f) n = 1e-12f; /* torch::nn::functional::normalize eps */
		for (int k = 0; k < 4; k++) rotations[4 * i + k] = q[k] / n;
	}
}
void orc_activations_backward(int P, const float* opacity_raw, const float* scaling_raw, const float* rotation_raw,
                              const float* dL_dopacity, const float* dL_dscales, const float* dL_drot,
                              float* g_opacity_raw, float* g_scaling_raw, float* g_rotation_raw)
{
	for (int i = 0; i < P; i++) {
		float s = 1.0f / (1.0f + expf(-opacity_raw[i]));
		g_opacity_raw[i] = dL_dopacity[i] * s * (1.0f - s);
		for (int k = 0; k < 3; k++) g_scaling_raw[3 * i + k] = dL_dscales[3 * i + k] * expf(scaling_raw[3 * i + k]);
		const float* q = rotation_raw + 4 * i; const float* d = dL_drot + 4 * i;
		float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
		if (n < 1e-12f) n = 1e-12f;
		float qd = (q[0] * d[0] + q[1] * d[1] + q[2] * d[2] + q[3] * d[3]) / (n * n);
		for (int k = 0; k < 4; k++) g_rotation_raw[4 * i + k] = (d[k] - q[k] * qd) / n;
	}
}

/* ---------------- simple-knn: third_party/simple-knn/simple_knn.cu:147-183 computes, per point, the mean of the
 * squared distances to its 3 nearest neighbours (exact search; Morton order and boxes only prune). ------------- */
void orc_knn_mean_dist2(int P, const float* pts, float* out)
{
#pragma omp parallel for schedule(static)
	for (int i = 0; i < P; i++) {
		float best[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
		for (int j = 0; j < P; j++) {
			if (j == i) continue;
			float dx = pts[3 * j] - pts[3 * i], dy = pts[3 * j + 1] - pts[3 * i + 1], dz = pts[3 * j + 2] - pts[3 * i + 2];
			float dist = dx * dx + dy * dy + dz * dz;
			for (int k = 0; k < 3; k++) if (best[k] > dist) { float t = best[k]; best[k] = dist; dist = t; }
		}
		out[i] = (best[0] + best[1] + best[2]) / 3.0f;
	}
}

/* ---------------- operate_points: src/operate_points.cu:38-71, cuda_rasterizer/operate_points.h:37-178 -------- */
void orc_transform_points(int P, const float* pts, const float* m, float* out)
{
	for (int i = 0; i < P; i++) transformPoint4x3(pts + 3 * i, m, out + 3 * i);
}
/* Quaternion through the 3x3 of m (Shoemake). Output order (w,x,y,z). NOTE: the reference's
 * insert_rot_to_rots (operate_points.h:170-178) writes z to slot +2 and never writes slot +3;
 * `reference_bug` != 0 reproduces that, 0 writes the mathematically intended (w,x,y,z). */
void orc_scale_transform_points(int P, float scale, const float* pts, const float* rots, const float* m,
                                const uint8_t* mask, float* out_pts, float* out_rots, int reference_bug)
{
	for (int i = 0; i < P; i++) {
		if (!mask[i]) continue;
		float p[3] = {pts[3 * i] * scale, pts[3 * i + 1] * scale, pts[3 * i + 2] * scale};
		transformPoint4x3(p, m, out_pts + 3 * i);
		float qx = rots[4 * i + 1], qy = rots[4 * i + 2], qz = rots[4 * i + 3], qw = rots[4 * i];
		float tx = 2.0f * qx, ty = 2.0f * qy, tz = 2.0f * qz;
		float twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
		float R0[3][3] = {{1.0f - (tyy + tzz), txy - twz, txz + twy}, {txy + twz, 1.0f - (txx + tzz), tyz - twx}, {txz - twy, tyz + twx, 1.0f - (txx + tyy)}};
		float R[3][3];
		for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[r][c] = m[r] * R0[0][c] + m[4 + r] * R0[1][c] + m[8 + r] * R0[2][c];
		float q[4]; /* x,y,z,w */
		float t = R[0][0] + R[1][1] + R[2][2];
		if (t > 0.0f) {
			t = sqrtf(t + 1.0f); q[3] = 0.5f * t; t = 0.5f / t;
			q[0] = (R[2][1] - R[1][2]) * t; q[1] = (R[0][2] - R[2][0]) * t; q[2] = (R[1][0] - R[0][1]) * t;
		} else {
			int a = 0; if (R[1][1] > R[0][0]) a = 1; if (R[2][2] > R[a][a]) a = 2;
			int b = (a + 1) % 3, c = (b + 1) % 3;
			t = sqrtf(R[a][a] - R[b][b] - R[c][c] + 1.0f);
			float xyz[3]; xyz[a] = 0.5f * t; t = 0.5f / t;
			q[3] = (R[c][b] - R[b][c]) * t; xyz[b] = (R[b][a] + R[a][b]) * t; xyz[c] = (R[c][a] + R[a][c]) * t;
			q[0] = xyz[0]; q[1] = xyz[1]; q[2] = xyz[2];
		}
		float* o = out_rots + 4 * i;
		o[0] = q[3]; o[1] = q[0];
		if (reference_bug) { o[2] = q[2]; } else { o[2] = q[1]; o[3] = q[2]; }
	}
}


/* ---------------- densification: GaussianModel::densifyAndPrune restated STEP BY STEP in the reference's own order
 * (src/gaussian_model.cpp:795-815 driver; :763-793 densifyAndClone; :716-761 densifyAndSplit; :644-714 densificationPostfix;
 * :588-642 prunePoints) — deliberately NOT the fused composition the CUDA kernel uses: each step re-materialises the model
 * like the reference's cat / mask-index chain, so the kernel's single-pass row bookkeeping is checked against an independent
 * derivation. The random draw at::normal(means, stds) (:734) is injected: samples = 0 + stds * z with z [2*n_split, 3]
 * supplied by the caller in the reference's repeat({N,1}) row order. Row widths: xyz 3 | f_dc 3 | f_rest 45 | opacity 1 |
 * scaling 3 | rotation 4. Pinned against the ATen ops the reference calls by tests/test_oracle_cpu.py (oracle/ref_densify.py). */
static const int DN_K[6] = {3, 3, 45, 1, 3, 4};
typedef struct { int n; float* t[18]; float *accum, *denom, *maxr; } dn_model;   /* t[0..5] params, [6..11] exp_avg, [12..17] exp_avg_sq */

static void dn_free(dn_model* a) { for (int i = 0; i < 18; i++) free(a->t[i]); free(a->accum); free(a->denom); free(a->maxr); }
static float dn_smax(const float* s3) { float a = expf(s3[0]), b = expf(s3[1]), c = expf(s3[2]); float m = a > b ? a : b; return m > c ? m : c; }

/* densificationPostfix (:644-714): parameters cat'ed, moments cat'ed with zeros, the three statistics RESET to zeros */
static void dn_postfix(dn_model* a, int n_new, float* const* ext /* 6 tensors of n_new rows */)
{
	for (int g = 0; g < 6; g++) {
		const int k = DN_K[g];
		for (int part = 0; part < 3; part++) {
			float* nt = (float*)calloc((size_t)(a->n + n_new) * k + 1, sizeof(float));
			memcpy(nt, a->t[6 * part + g], (size_t)a->n * k * sizeof(float));
			if (part == 0 && n_new) memcpy(nt + (size_t)a->n * k, ext[g], (size_t)n_new * k * sizeof(float));
			free(a->t[6 * part + g]);
			a->t[6 * part + g] = nt;
		}
	}
	a->n += n_new;
	free(a->accum); free(a->denom); free(a->maxr);
	a->accum = (float*)calloc((size_t)a->n + 1, sizeof(float));
	a->denom = (float*)calloc((size_t)a->n + 1, sizeof(float));
	a->maxr = (float*)calloc((size_t)a->n + 1, sizeof(float));
}
/* prunePoints (:588-642): every tensor, both moments and the statistics indexed by ~mask */
static void dn_prune(dn_model* a, const uint8_t* mask)
{
	int keep = 0;
	for (int i = 0; i < a->n; i++) keep += !mask[i];
	for (int ti = 0; ti < 18; ti++) {
		const int k = DN_K[ti % 6];
		float* nt = (float*)calloc((size_t)keep * k + 1, sizeof(float));
		int d = 0;
		for (int i = 0; i < a->n; i++) if (!mask[i]) { memcpy(nt + (size_t)d * k, a->t[ti] + (size_t)i * k, k * sizeof(float)); d++; }
		free(a->t[ti]); a->t[ti] = nt;
	}
	float* st[3] = {a->accum, a->denom, a->maxr};
	for (int q = 0; q < 3; q++) { int d = 0; for (int i = 0; i < a->n; i++) if (!mask[i]) st[q][d++] = st[q][i]; }
	a->n = keep;
}

/* number of rows densifyAndSplit selects (the caller needs it to draw z [2*n_split, 3]) */
int orc_densify_split_count(int P, const float* scaling, const float* accum, const float* denom, float max_grad, float extent, float percent_dense)
{
	int n = 0;
	for (int i = 0; i < P; i++) {
		float g = accum[i] / denom[i];
		if (isnan(g)) g = 0.0f;
		if (g >= max_grad && dn_smax(scaling + 3 * i) > percent_dense * extent) n++;
	}
	return n;
}

/* in: 18 tensors of P rows + statistics; out: 18 tensors with room for 2*P rows each; returns the new row count */
int orc_densify_and_prune(int P, const float* const* in18, const float* accum, const float* denom, const float* max_radii2D,
                          float max_grad, float min_opacity, float extent, float percent_dense, int max_screen_size,
                          const float* z, float* const* out18)
{
	dn_model a;
	a.n = P;
	for (int ti = 0; ti < 18; ti++) {
		const size_t bytes = (size_t)P * DN_K[ti % 6] * sizeof(float);
		a.t[ti] = (float*)malloc(bytes + 4); memcpy(a.t[ti], in18[ti], bytes);
	}
	a.accum = (float*)malloc((size_t)P * 4 + 4); memcpy(a.accum, accum, (size_t)P * 4);
	a.denom = (float*)malloc((size_t)P * 4 + 4); memcpy(a.denom, denom, (size_t)P * 4);
	a.maxr = (float*)malloc((size_t)P * 4 + 4); memcpy(a.maxr, max_radii2D, (size_t)P * 4);
	/* :801-802  grads = xyz_gradient_accum / denom; grads[isnan] = 0 */
	float* grads = (float*)malloc((size_t)P * 4 + 4);
	for (int i = 0; i < P; i++) { float g = a.accum[i] / a.denom[i]; grads[i] = isnan(g) ? 0.0f : g; }

	/* ---- densifyAndClone (:763-793) */
	{
		uint8_t* sel = (uint8_t*)calloc((size_t)a.n + 1, 1);
		int n_new = 0;
		for (int i = 0; i < a.n; i++) {
			sel[i] = sqrtf(grads[i] * grads[i]) >= max_grad && dn_smax(a.t[4] + 3 * i) <= percent_dense * extent;  /* frobenius_norm over the size-1 dim */
			n_new += sel[i];
		}
		float* ext[6];
		for (int g = 0; g < 6; g++) {
			ext[g] = (float*)malloc((size_t)n_new * DN_K[g] * 4 + 4);
			int d = 0;
			for (int i = 0; i < a.n; i++) if (sel[i]) { memcpy(ext[g] + (size_t)d * DN_K[g], a.t[g] + (size_t)i * DN_K[g], DN_K[g] * 4); d++; }
		}
		dn_postfix(&a, n_new, ext);
		for (int g = 0; g < 6; g++) free(ext[g]);
		free(sel);
	}
	/* ---- densifyAndSplit (:716-761), N = 2 */
	{
		const int N = 2, n_init = a.n;
		uint8_t* sel = (uint8_t*)calloc((size_t)n_init + 1, 1);
		int ns = 0;
		for (int i = 0; i < n_init; i++) {
			const float pg = i < P ? grads[i] : 0.0f;     /* padded_grad: zeros beyond grads.size(0) */
			sel[i] = pg >= max_grad && dn_smax(a.t[4] + 3 * i) > percent_dense * extent;
			ns += sel[i];
		}
		float* ext[6];
		for (int g = 0; g < 6; g++) ext[g] = (float*)malloc((size_t)N * ns * DN_K[g] * 4 + 4);
		const float inv = 1.0f / (float)(0.8 * N);        /* ATen divides by a CPU scalar through its reciprocal */
		for (int c = 0; c < N; c++) {
			int d = 0;
			for (int i = 0; i < n_init; i++) if (sel[i]) {
				const int j = c * ns + d;                 /* repeat({N,1}): copy-major */
				const float* q = a.t[5] + 4 * i;          /* build_rotation (include/general_utils.h:31-56) on the RAW rotation */
				const float nr = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
				const float r = q[0] / nr, x = q[1] / nr, y = q[2] / nr, zq = q[3] / nr;
				const float R[9] = {1 - 2 * (y * y + zq * zq), 2 * (x * y - r * zq), 2 * (x * zq + r * y),
				                    2 * (x * y + r * zq), 1 - 2 * (x * x + zq * zq), 2 * (y * zq - r * x),
				                    2 * (x * zq - r * y), 2 * (y * zq + r * x), 1 - 2 * (x * x + y * y)};
				float smp[3];
				for (int k = 0; k < 3; k++) smp[k] = 0.0f + expf(a.t[4][3 * i + k]) * z[3 * j + k];   /* at::normal(means, stds) with injected z */
				for (int k = 0; k < 3; k++)
					ext[0][3 * j + k] = R[3 * k] * smp[0] + R[3 * k + 1] * smp[1] + R[3 * k + 2] * smp[2] + a.t[0][3 * i + k];   /* bmm + xyz */
				for (int k = 0; k < 3; k++) ext[4][3 * j + k] = logf(expf(a.t[4][3 * i + k]) * inv);
				memcpy(ext[1] + 3 * j, a.t[1] + 3 * i, 12);
				memcpy(ext[2] + 45 * (size_t)j, a.t[2] + 45 * (size_t)i, 180);
				ext[3][j] = a.t[3][i];
				memcpy(ext[5] + 4 * j, a.t[5] + 4 * i, 16);
				d++;
			}
		}
		dn_postfix(&a, N * ns, ext);
		for (int g = 0; g < 6; g++) free(ext[g]);
		uint8_t* filt = (uint8_t*)calloc((size_t)a.n + 1, 1);   /* cat(selected_pts_mask, zeros(N * n_sel)) */
		memcpy(filt, sel, n_init);
		dn_prune(&a, filt);
		free(filt); free(sel);
	}
	/* ---- final prune (:806-812) */
	{
		uint8_t* mask = (uint8_t*)calloc((size_t)a.n + 1, 1);
		for (int i = 0; i < a.n; i++) {
			mask[i] = (1.0f / (1.0f + expf(-a.t[3][i]))) < min_opacity;
			if (max_screen_size) {
				const int big_vs = a.maxr[i] > (float)max_screen_size;          /* never true: postfix zeroed max_radii2D_ (:711) */
				const int big_ws = dn_smax(a.t[4] + 3 * i) > 0.1f * extent;
				mask[i] = mask[i] || big_vs || big_ws;
			}
		}
		dn_prune(&a, mask);
		free(mask);
	}
	for (int ti = 0; ti < 18; ti++) memcpy(out18[ti], a.t[ti], (size_t)a.n * DN_K[ti % 6] * sizeof(float));
	const int n = a.n;
	free(grads);
	dn_free(&a);
	return n;
}

/* resetOpacity (:556-565): inverse_sigmoid(min(sigmoid(o), ones_like(sigmoid(o) * 0.01))) — the clamp is a no-op (quirk 9) */
void orc_reset_opacity(int P, float* opacity)
{
	for (int i = 0; i < P; i++) {
		float a = 1.0f / (1.0f + expf(-opacity[i]));
		if (a > 1.0f) a = 1.0f;
		opacity[i] = logf(a / (1.0f - a));
	}
}

/* ---------------- Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) — the counter-based
 * generator the fused densify kernel draws its split samples from (csrc/psb_densify.cu:philox4x32_10, same round structure and constants).
 * Pinned by the Random123 known-answer vectors in tests/test_oracle_cpu.py. ctr[4], key[2] -> out[4]. */
void orc_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out)
{
	const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
	uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
	for (int r = 0; r < 10; r++) {
		const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
		const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
		const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += W0; k1 += W1;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

int orc_version(void) { return 3; }
