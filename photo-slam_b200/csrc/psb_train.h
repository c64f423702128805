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

// ---- NVLink data-parallel step (psb_dp_*, include/psb200.h) -----------------------------------------------------------
// Ownership: Gaussians are cut into chunks of 128 (one block of the per-Gaussian backward); chunk c belongs to rank
// c % world and is that rank's local chunk c / world. Every rank holds a symmetric arena (same layout, peer-mapped through
// CUDA IPC): the six parameter tensors, an inbox [world][nlocal_max][128][20 floats], per-rank meta rows and epoch flags.
constexpr int DP_MAX_WORLD = 8;
constexpr int DP_REC = 20;   // floats per gradient record: g_xyz 3 | g_dc 3 | g_opacity 1 | g_scaling 3 | g_rotation 4 | masked dL/dRGB 3 | pad 2 | epoch
struct DpPush {              // what the per-Gaussian backward needs to deliver its records
	float* inbox[DP_MAX_WORLD];        // inbox base of every rank (own included), peer-mapped
	float* meta[DP_MAX_WORLD];         // [world][8] floats on every rank: row r = camera centre + SH degree of rank r this step
	uint32_t* grad_flag[DP_MAX_WORLD]; // [world] epoch words on every rank: word r = "rank r's records of this epoch have landed"
	uint32_t* done_counter;            // local
	int world, rank, nlocal_max;
	uint32_t epoch;
	int fence_in_kernel;               // 1: every block fences its remote stores, the grid's last block raises the flags (PSB_DP_SIGNAL=fence)
	                                   // 0: a one-block signal kernel behind the data kernel does (stream order + ONE system fence; default)
};
struct DpShard {             // what the owner-side Adam kernels need
	float* param[DP_MAX_WORLD][6];     // the six parameter tensors on every rank (row of `rank` = local)
	uint32_t* param_flag[DP_MAX_WORLD];// [world] epoch words on every rank: word r = "rank r's updated rows have landed"
	const float* inbox;                // local inbox
	const float* meta;                 // local meta rows
	float* g_rest;                     // local scratch [nlocal_max*128][45]: summed f_rest gradient of the owned rows
	uint32_t* done_counter;            // local
	int world, rank, nlocal_max, nlocal, P;   // nlocal: owned chunks handled by this launch, starting at local chunk lc_first
	int lc_first;
	int bulk;                          // 1 (PSB_DP_BULK=1): the small-parameter chunks leave as TMA bulk stores; 0 (default): per-thread 16-byte stores
	int rotate;                        // 1: each rank walks the destination ranks starting at rank + 1 (PSB_DP_ROTATE=0: all start at rank 0)
	uint32_t epoch;
	int fence_in_kernel;               // as in DpPush
};
int launch_push_backward(int first, int P, const TrainTensors& t, const Camera& cam, const GeomState& geom, float* sink, const StepHyper& h,
                         const DensifyStats& st, const uint32_t* counters, uint32_t capacity, const DpPush& dp, cudaStream_t stream);
int launch_shard_adam(const DpShard& d, const TrainTensors& t, const StepHyper& h, float grad_scale, cudaStream_t stream, cudaEvent_t between = nullptr);
// spins (device side) until flags[0..world) >= epoch (world <= 64 words); after `timeout_ms` writes 1 to *status and gives up
int launch_wait_flags(const uint32_t* flags, int world, uint32_t epoch, uint32_t* status, cudaStream_t stream);

// Gaussians [first, P) (first must be a multiple of 128 so f_rest chunks stay 16-byte aligned)
// seeds: scratch [P][20] floats (per-Gaussian SH gradient seeds handed from the per-Gaussian kernel to the f_rest stream kernel)
int launch_fused_backward(bool adam, int first, int P, const TrainTensors& t, const Camera& cam, const GeomState& geom, float* sink, float* seeds,
                          const StepHyper& h, const GradSegments& grads, const DensifyStats& st, const uint32_t* counters, uint32_t capacity,
                          cudaStream_t stream, cudaEvent_t between = nullptr /* recorded between the two kernels (stage timing) */);
int launch_adam(size_t n, float* p, float* m, float* v, const float* g, float lr, const StepHyper& h, float grad_scale, cudaStream_t stream);
int launch_loss(int H, int W, const float* img, const float* gt, const float* mask, float lambda_dssim, float* dmap, double* sums,
                float* dL_dimg, cudaStream_t stream);

}  // namespace psb
