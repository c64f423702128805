// Internal interface of the trainer-step kernels (psb_train.cu, psb_loss.cu) — not installed.
#pragma once
#include "psb_kernels.h"

namespace psb {

// Raw trainer tensors in the reference's layout and order (include/gaussian_model.h; optimizer groups of
// src/gaussian_model.cpp:477-503): 0 xyz [P,3] | 1 features_dc [P,1,3] | 2 features_rest [P,15,3] |
// 3 opacity [P,1] | 4 scaling [P,3] | 5 rotation [P,4].  m / v = Adam exp_avg / exp_avg_sq of the same shapes.
struct TrainTensors {
	float* p[6];
	float* m[6];
	float* v[6];
};
struct GradSegments {
	float* g[6];
};
struct StepHyper {
	float lr[6];
	float beta1, beta2, eps;
	float inv_bc1;   // 1 / (1 - beta1^t)
	float bc2_sqrt;  // sqrt(1 - beta2^t)
	int D;           // active SH degree
};
struct DensifyStats {
	int enabled;
	float* max_radii2D;         // [P]
	float* xyz_gradient_accum;  // [P,1]
	float* denom;               // [P,1]
};

// Gaussians [first, P) (first must be a multiple of 128 so f_rest chunks stay 16-byte aligned)
// seeds: scratch [P][20] floats (per-Gaussian SH gradient seeds handed from the per-Gaussian kernel to the f_rest stream kernel)
int launch_fused_backward(bool adam, int first, int P, const TrainTensors& t, const Camera& cam, const GeomState& geom, float* sink, float* seeds,
                          const StepHyper& h, const GradSegments& grads, const DensifyStats& st, const uint32_t* counters, uint32_t capacity,
                          cudaStream_t stream);
int launch_adam(size_t n, float* p, float* m, float* v, const float* g, float lr, const StepHyper& h, float grad_scale, cudaStream_t stream);
int launch_loss(int H, int W, const float* img, const float* gt, const float* mask, float lambda_dssim, float* dmap, double* sums,
                float* dL_dimg, cudaStream_t stream);

}  // namespace psb
