// Internal kernel-launch interface shared by the translation units of libpsb200 (not installed).
#pragma once
#include "psb_common.cuh"
#include "psb_state.h"

namespace psb {

// Per-Gaussian inputs of the forward/backward preprocess. Exactly one of {shs | sh_dc+sh_rest | colors_precomp}
// and one of {scales+rotations | cov3D_precomp} is non-null ("None" = nullptr, the reference's empty-tensor
// convention, reference gaussian_rasterizer.cpp:209-219).
struct GaussIn {
	int P, D, M;
	const float* means3D;        // [P,3]
	const float* scales;         // [P,3]
	const float* rotations;      // [P,4] (w,x,y,z)
	const float* opacities;      // [P]
	const float* shs;            // [P,M,3]
	const float* sh_dc;          // [P,1,3]   (raw trainer layout)
	const float* sh_rest;        // [P,M-1,3] (raw trainer layout)
	const float* cov3D_precomp;  // [P,6]
	const float* colors_precomp; // [P,3]
	float scale_modifier;
	int sh_vec4;                 // shs rows are 16-byte aligned and M == 16: use 128-bit loads
};

// Where the tile backward kernel adds its 9 per-Gaussian sums. Strides in floats.
struct GradSink {
	float* mean2D;  int mean2D_stride;   // .x .y at +0 +1
	float* conic;   int conic_stride;    // .x .y .w at +0 +1 +3
	float* opacity; int opacity_stride;
	float* color;   int color_stride;    // rgb at +0..2
	int packed;                          // != 0: one 16-byte aligned [P][12] row layout (mean2D base), vector reductions
};

// Outputs of the per-Gaussian backward (any pointer may be null = not wanted).
struct GaussGradOut {
	float* dL_dmeans3D;  // [P,3]
	float* dL_dcov3D;    // [P,6]
	float* dL_dsh;       // [P,M,3]
	float* dL_dscales;   // [P,3]
	float* dL_drots;     // [P,4]
};

// tight: instance lists hold only the tiles a splat can reach (trainer path; see preprocess_fwd_kernel) — pass the same flag to launch_scan_binning
int launch_preprocess(const GaussIn& in, const Camera& cam, int* radii_out, const GeomState& geom, bool raw, bool tight, cudaStream_t stream);
int launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present, cudaStream_t stream);
int launch_depth_sort_and_scan(int P, GeomState& geom, bool scan, cudaStream_t stream);
int launch_binning(int P, const Camera& cam, const GeomState& geom, BinState& bin, const ImgState& img, size_t n_host, cudaStream_t stream);
int launch_scan_binning(int P, const Camera& cam, const GeomState& geom, BinState& bin, const ImgState& img, size_t capacity, bool tight,
                        uint32_t* ovf, uint32_t seq, cudaStream_t stream);
int launch_render_forward(const Camera& cam, const uint2* ranges, const uint32_t* point_list, const GaussRec* rec, const float* bg,
                          float* out_color, float* final_T, uint32_t* n_contrib, cudaStream_t stream);
int launch_render_backward(const Camera& cam, const uint2* ranges, const uint32_t* point_list, const GaussRec* rec, const float* bg,
                           const float* final_T, const uint32_t* n_contrib, const float* dL_dpix, const GradSink& sink,
                           cudaStream_t stream);
int launch_preprocess_backward(const GaussIn& in, const Camera& cam, const GeomState& geom, const float* dL_dmean2D, int mean2D_stride,
                               const float* dL_dconic, int conic_stride, const float* dL_dcolor, int color_stride,
                               const GaussGradOut& out, cudaStream_t stream);

}  // namespace psb
