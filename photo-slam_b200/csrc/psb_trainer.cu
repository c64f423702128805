// Trainer-level C-ABI (declared in include/psb200.h): a context that owns the scratch arena of the fused
// training iteration, and the iteration itself:
//   preprocess(raw params) -> depth sort -> scan -> emit -> tile sort -> ranges -> render
//   -> fused L1+SSIM loss fwd/bwd -> tile backward -> fused per-Gaussian backward + Adam (+ densify stats)
// No host synchronisation inside a step: the instance count stays on the device (the binning arena is sized
// from the previous iterations and grown on demand; an overflowing step degrades to a no-op and is reported
// by psb_trainer_result so the caller repeats it).
#include <cmath>
#include <cstring>
#include <cstdlib>
#include "psb_train.h"
#include "../../include/psb200.h"

using namespace psb;

struct psb_trainer {
	char* geom_chunk = nullptr; size_t geom_bytes = 0; int geom_P = -1;
	char* img_chunk = nullptr; size_t img_bytes = 0; int img_N = -1;
	char* bin_chunk = nullptr; size_t bin_bytes = 0; size_t capacity = 0;
	float* image = nullptr; float* dL_dpix = nullptr; float* dmap = nullptr; size_t pix_alloc = 0;
	float* sink = nullptr; float* seeds = nullptr; size_t sink_P = 0;
	double* sums = nullptr;      // device [2]
	double* h_sums = nullptr;    // pinned [2]
	uint32_t* h_count = nullptr; // pinned [1]
	cudaEvent_t readback = nullptr;  // recorded once the instance count (and the loss sums) of the last call are in pinned memory
	int last_P = 0, last_W = 0, last_H = 0;
	float last_lambda = 0.2f;
	bool have_loss = false;
	// tight instance lists: only the tiles a splat can reach get an instance (see preprocess_fwd_kernel); PSB_TIGHT=0 keeps the
	// reference's full rectangles (A/B measurements only — images and gradients are the same either way)
	bool tight = !(getenv("PSB_TIGHT") && atoi(getenv("PSB_TIGHT")) == 0);
	// optional per-stage timing (CUDA events on the step's stream)
	bool profiling = false;
	static constexpr int NSTAGE = 8;
	cudaEvent_t ev[NSTAGE + 1] = {};
	bool ev_ready = false, ev_recorded = false;
	void mark(int i, cudaStream_t s) { if (profiling && ev_ready) cudaEventRecord(ev[i], s); }
};

namespace {

template <typename T>
int dev_alloc(T** p, size_t count)
{
	if (*p) { cudaFree(*p); *p = nullptr; }
	PSB_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T) + 256));
	return 0;
}

int ensure(psb_trainer* t, int P, int W, int H, cudaStream_t stream)
{
	const size_t N = (size_t)W * H;
	int rc;
	if (t->geom_P != P) {
		PSB_CUDA_OK(cudaStreamSynchronize(stream));
		t->geom_bytes = required_bytes<GeomState>((size_t)P);
		if ((rc = dev_alloc(&t->geom_chunk, t->geom_bytes))) return rc;
		t->geom_P = P;
	}
	if (t->img_N != (int)N) {
		PSB_CUDA_OK(cudaStreamSynchronize(stream));
		t->img_bytes = required_bytes<ImgState>(N);
		if ((rc = dev_alloc(&t->img_chunk, t->img_bytes))) return rc;
		t->img_N = (int)N;
	}
	if (t->pix_alloc < N) {
		PSB_CUDA_OK(cudaStreamSynchronize(stream));
		if ((rc = dev_alloc(&t->image, 3 * N))) return rc;
		if ((rc = dev_alloc(&t->dL_dpix, 3 * N))) return rc;
		if ((rc = dev_alloc(&t->dmap, 9 * N))) return rc;
		t->pix_alloc = N;
	}
	if (t->sink_P < (size_t)P) {
		PSB_CUDA_OK(cudaStreamSynchronize(stream));
		if ((rc = dev_alloc(&t->sink, (size_t)P * 12))) return rc;
		if ((rc = dev_alloc(&t->seeds, (size_t)P * 20))) return rc;
		PSB_CUDA_OK(cudaMemsetAsync(t->sink, 0, (size_t)P * 12 * sizeof(float), stream));
		t->sink_P = (size_t)P;
	}
	if (!t->sums) {
		if ((rc = dev_alloc(&t->sums, 2))) return rc;
		PSB_CUDA_OK(cudaMallocHost(reinterpret_cast<void**>(&t->h_sums), 2 * sizeof(double)));
		PSB_CUDA_OK(cudaMallocHost(reinterpret_cast<void**>(&t->h_count), sizeof(uint32_t)));
		PSB_CUDA_OK(cudaEventCreateWithFlags(&t->readback, cudaEventDisableTiming));
	}
	const size_t want = (size_t)P * 6 + (1u << 16);
	if (t->capacity == 0 || (t->capacity < want && t->last_P != P)) {
		PSB_CUDA_OK(cudaStreamSynchronize(stream));
		t->capacity = t->capacity > want ? t->capacity : want;
		t->bin_bytes = required_bytes<BinState>(t->capacity);
		if ((rc = dev_alloc(&t->bin_chunk, t->bin_bytes))) return rc;
	}
	t->last_P = P; t->last_W = W; t->last_H = H;
	return 0;
}

Camera to_camera(const psb_camera* c)
{
	Camera cam;
	cam.view = c->viewmatrix; cam.proj = c->projmatrix; cam.campos = c->campos;
	cam.tan_fovx = c->tan_fovx; cam.tan_fovy = c->tan_fovy;
	cam.focal_y = c->height / (2.0f * c->tan_fovy);
	cam.focal_x = c->width / (2.0f * c->tan_fovx);
	cam.W = c->width; cam.H = c->height;
	cam.grid_x = (c->width + PSB_TILE_X - 1) / PSB_TILE_X;
	cam.grid_y = (c->height + PSB_TILE_Y - 1) / PSB_TILE_Y;
	return cam;
}

GaussIn raw_input(int P, int D, int M, const psb_model* m)
{
	GaussIn in;
	memset(&in, 0, sizeof(in));
	in.P = P; in.D = D; in.M = M;
	in.means3D = m->param[0]; in.sh_dc = m->param[1]; in.sh_rest = m->param[2];
	in.opacities = m->param[3]; in.scales = m->param[4]; in.rotations = m->param[5];
	in.scale_modifier = 1.0f;
	return in;
}

StepHyper to_hyper(const psb_step* s)
{
	StepHyper h;
	for (int i = 0; i < 6; i++) h.lr[i] = s->lr[i];
	h.beta1 = s->beta1; h.beta2 = s->beta2; h.eps = s->eps;
	h.inv_bc1 = (float)(1.0 / (1.0 - std::pow((double)s->beta1, s->step)));
	h.bc2_sqrt = (float)std::sqrt(1.0 - std::pow((double)s->beta2, s->step));
	h.D = s->sh_degree;
	return h;
}

int check_model(int P, int M, const psb_model* m, bool need_moments)
{
	if (P < 0 || M != 16 || !m) { set_error_msg("psb_trainer: need P >= 0, M == 16 (SH degree 3 storage) and a model"); return PSB_ERR_ARG; }
	for (int i = 0; i < 6; i++) {
		if (P > 0 && (!m->param[i] || (need_moments && (!m->exp_avg[i] || !m->exp_avg_sq[i])))) { set_error_msg("psb_trainer: null model tensor"); return PSB_ERR_ARG; }
	}
	const uintptr_t a = reinterpret_cast<uintptr_t>(m->param[2]) | reinterpret_cast<uintptr_t>(m->param[5]) |
	                    (need_moments ? (reinterpret_cast<uintptr_t>(m->exp_avg[2]) | reinterpret_cast<uintptr_t>(m->exp_avg_sq[2]) |
	                                     reinterpret_cast<uintptr_t>(m->exp_avg[5]) | reinterpret_cast<uintptr_t>(m->exp_avg_sq[5])) : 0);
	if (a & 15) { set_error_msg("psb_trainer: features_rest / rotation tensors must be 16-byte aligned"); return PSB_ERR_ARG; }
	return 0;
}

// forward from raw parameters into t->image (or out_color)
int forward_raw(psb_trainer* t, int P, int M, int D, const psb_model* model, const psb_camera* camera, const float* background,
                float* out_color, int* radii, GeomState& geom, BinState& bin, ImgState& img, Camera& cam, cudaStream_t stream)
{
	int rc;
	if ((rc = ensure(t, P, camera->width, camera->height, stream))) return rc;
	cam = to_camera(camera);
	char* gc = t->geom_chunk; geom = GeomState::from_chunk(gc, (size_t)P);
	char* ic = t->img_chunk; img = ImgState::from_chunk(ic, (size_t)camera->width * camera->height);
	char* bc = t->bin_chunk; bin = BinState::from_chunk(bc, t->capacity);
	const GaussIn in = raw_input(P, D, M, model);
	t->mark(0, stream);
	if ((rc = launch_preprocess(in, cam, radii, geom, /*raw=*/true, t->tight, stream))) return rc;
	t->mark(1, stream);
	if ((rc = launch_depth_sort_and_scan(P, geom, /*scan=*/false, stream))) return rc;
	t->mark(2, stream);
	if ((rc = launch_scan_binning(P, cam, geom, bin, img, t->capacity, t->tight, stream))) return rc;
	t->mark(3, stream);
	// the instance count travels to pinned memory as soon as it exists: psb_trainer_result waits on `readback`, not on the stream
	PSB_CUDA_OK(cudaMemcpyAsync(t->h_count, geom.counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
	PSB_CUDA_OK(cudaEventRecord(t->readback, stream));
	const int res = make_sort_plan(tile_id_bits(cam.grid_x * cam.grid_y)).npass & 1;
	rc = launch_render_forward(cam, img.ranges, bin.inst[res], geom.rec, background, out_color ? out_color : t->image, img.final_T,
	                           img.n_contrib, stream);
	t->mark(4, stream);
	return rc;
}

int step_impl(psb_trainer* t, int P, int M, const psb_model* model, const psb_camera* camera, const float* background, const float* gt_image,
              const float* mask, const psb_step* step, float* out_color, int* radii, float* const* grads, cudaStream_t stream,
              bool tiles_only = false)
{
	int rc;
	if (!t || !camera || !step || !background || !gt_image) { set_error_msg("psb_trainer_step: null argument"); return PSB_ERR_ARG; }
	if ((rc = check_model(P, M, model, grads == nullptr && !tiles_only))) return rc;
	if (step->sh_degree < 0 || step->sh_degree > 3 || step->step < 1) { set_error_msg("psb_trainer_step: sh_degree in 0..3 and step >= 1 required"); return PSB_ERR_ARG; }
	GeomState geom; BinState bin; ImgState img; Camera cam;
	if ((rc = forward_raw(t, P, M, step->sh_degree, model, camera, background, out_color, radii, geom, bin, img, cam, stream))) return rc;
	const float* image = out_color ? out_color : t->image;
	if ((rc = launch_loss(cam.H, cam.W, image, gt_image, mask, step->lambda_dssim, t->dmap, t->sums, t->dL_dpix, stream))) return rc;
	// loss sums -> pinned memory right behind the loss kernel, so the host can read the loss of this iteration (and start
	// enqueueing the next one) while the backward half is still running
	PSB_CUDA_OK(cudaMemcpyAsync(t->h_sums, t->sums, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
	PSB_CUDA_OK(cudaEventRecord(t->readback, stream));
	t->last_lambda = step->lambda_dssim;
	t->have_loss = true;
	t->mark(5, stream);
	if (P == 0) return 0;
	GradSink sink;
	sink.mean2D = t->sink; sink.mean2D_stride = 12;
	sink.conic = t->sink + 3; sink.conic_stride = 12;
	sink.opacity = t->sink + 7; sink.opacity_stride = 12;
	sink.color = t->sink + 8; sink.color_stride = 12;
	sink.packed = 1;
	const int res = make_sort_plan(tile_id_bits(cam.grid_x * cam.grid_y)).npass & 1;
	if ((rc = launch_render_backward(cam, img.ranges, bin.inst[res], geom.rec, background, img.final_T, img.n_contrib, t->dL_dpix, sink, stream))) return rc;
	t->mark(6, stream);
	if (tiles_only) return 0;
	TrainTensors tt;
	for (int i = 0; i < 6; i++) { tt.p[i] = model->param[i]; tt.m[i] = model->exp_avg[i]; tt.v[i] = model->exp_avg_sq[i]; }
	GradSegments gs;
	for (int i = 0; i < 6; i++) gs.g[i] = grads ? grads[i] : nullptr;
	DensifyStats st;
	st.enabled = (step->update_densify_stats && model->max_radii2D && model->xyz_gradient_accum && model->denom) ? 1 : 0;
	st.max_radii2D = model->max_radii2D; st.xyz_gradient_accum = model->xyz_gradient_accum; st.denom = model->denom;
	const StepHyper h = to_hyper(step);
	rc = launch_fused_backward(grads == nullptr, 0, P, tt, cam, geom, t->sink, t->seeds, h, gs, st, geom.counters, (uint32_t)t->capacity, stream);
	t->mark(7, stream);
	t->ev_recorded = t->profiling && t->ev_ready;
	return rc;
}

}  // namespace

extern "C" {

int psb_trainer_create(psb_trainer** out)
{
	if (!out) { set_error_msg("psb_trainer_create: null"); return PSB_ERR_ARG; }
	*out = new psb_trainer();
	return 0;
}

int psb_trainer_destroy(psb_trainer* t)
{
	if (!t) return 0;
	cudaFree(t->geom_chunk); cudaFree(t->img_chunk); cudaFree(t->bin_chunk); cudaFree(t->image); cudaFree(t->dL_dpix); cudaFree(t->dmap);
	cudaFree(t->sink); cudaFree(t->seeds); cudaFree(t->sums);
	if (t->h_sums) cudaFreeHost(t->h_sums);
	if (t->h_count) cudaFreeHost(t->h_count);
	if (t->readback) cudaEventDestroy(t->readback);
	delete t;
	return 0;
}

int psb_trainer_render(psb_trainer* t, int P, int M, const psb_model* model, const psb_camera* camera, const float* background, int sh_degree,
                       float* out_color, int* radii, void* stream_)
{
	int rc;
	if (!t || !camera || !background || !out_color) { set_error_msg("psb_trainer_render: null argument"); return PSB_ERR_ARG; }
	if ((rc = check_model(P, M, model, false))) return rc;
	GeomState geom; BinState bin; ImgState img; Camera cam;
	t->have_loss = false;
	return forward_raw(t, P, M, sh_degree, model, camera, background, out_color, radii, geom, bin, img, cam, (cudaStream_t)stream_);
}

int psb_trainer_step(psb_trainer* t, int P, int M, const psb_model* model, const psb_camera* camera, const float* background,
                     const float* gt_image, const float* mask, const psb_step* step, float* out_color, int* radii, void* stream_)
{
	return step_impl(t, P, M, model, camera, background, gt_image, mask, step, out_color, radii, nullptr, (cudaStream_t)stream_);
}

int psb_trainer_backward(psb_trainer* t, int P, int M, const psb_model* model, const psb_camera* camera, const float* background,
                         const float* gt_image, const float* mask, const psb_step* step, float* out_color, int* radii, float* const* grads,
                         void* stream_)
{
	if (!grads) { set_error_msg("psb_trainer_backward: grads required"); return PSB_ERR_ARG; }
	for (int i = 0; i < 6; i++) if (P > 0 && !grads[i]) { set_error_msg("psb_trainer_backward: null gradient segment"); return PSB_ERR_ARG; }
	return step_impl(t, P, M, model, camera, background, gt_image, mask, step, out_color, radii, grads, (cudaStream_t)stream_);
}

int psb_trainer_backward_begin(psb_trainer* t, int P, int M, const psb_model* model, const psb_camera* camera, const float* background,
                               const float* gt_image, const float* mask, const psb_step* step, float* out_color, int* radii, void* stream_)
{
	return step_impl(t, P, M, model, camera, background, gt_image, mask, step, out_color, radii, nullptr, (cudaStream_t)stream_, /*tiles_only=*/true);
}

int psb_trainer_backward_slab(psb_trainer* t, int P, int M, const psb_model* model, const psb_camera* camera, const psb_step* step, int first,
                              int count, float* const* grads, void* stream_)
{
	int rc;
	if (!t || !camera || !step || !grads || t->geom_P != P) { set_error_msg("psb_trainer_backward_slab: bad argument / no matching backward_begin"); return PSB_ERR_ARG; }
	if ((rc = check_model(P, M, model, false))) return rc;
	if (first < 0 || count < 0 || first + count > P || (first % 128) != 0) { set_error_msg("psb_trainer_backward_slab: slab must start at a multiple of 128"); return PSB_ERR_ARG; }
	if (count == 0) return 0;
	char* gc = t->geom_chunk;
	GeomState geom = GeomState::from_chunk(gc, (size_t)P);
	const Camera cam = to_camera(camera);
	TrainTensors tt;
	for (int i = 0; i < 6; i++) { tt.p[i] = model->param[i]; tt.m[i] = nullptr; tt.v[i] = nullptr; }
	GradSegments gs;
	for (int i = 0; i < 6; i++) gs.g[i] = grads[i];
	DensifyStats st;
	st.enabled = (step->update_densify_stats && model->max_radii2D && model->xyz_gradient_accum && model->denom) ? 1 : 0;
	st.max_radii2D = model->max_radii2D; st.xyz_gradient_accum = model->xyz_gradient_accum; st.denom = model->denom;
	const StepHyper h = to_hyper(step);
	return launch_fused_backward(false, first, first + count, tt, cam, geom, t->sink, t->seeds, h, gs, st, geom.counters, (uint32_t)t->capacity, (cudaStream_t)stream_);
}

int psb_adam_update(int P, int M, const psb_model* model, float* const* grads, const psb_step* step, float grad_scale, void* stream_)
{
	int rc;
	if ((rc = check_model(P, M, model, true))) return rc;
	if (!grads || !step || step->step < 1) { set_error_msg("psb_adam_update: bad argument"); return PSB_ERR_ARG; }
	const StepHyper h = to_hyper(step);
	const size_t per[6] = {3, 3, (size_t)(M - 1) * 3, 1, 3, 4};
	for (int i = 0; i < 6; i++)
		if ((rc = launch_adam((size_t)P * per[i], model->param[i], model->exp_avg[i], model->exp_avg_sq[i], grads[i], step->lr[i], h, grad_scale,
		                      (cudaStream_t)stream_)))
			return rc;
	return 0;
}

int psb_adam_flat(size_t n, float* param, float* exp_avg, float* exp_avg_sq, const float* grad, float lr, const psb_step* step,
                  float grad_scale, void* stream_)
{
	if (!step || step->step < 1 || (n > 0 && (!param || !exp_avg || !exp_avg_sq || !grad))) { set_error_msg("psb_adam_flat: bad argument"); return PSB_ERR_ARG; }
	const StepHyper h = to_hyper(step);
	return launch_adam(n, param, exp_avg, exp_avg_sq, grad, lr, h, grad_scale, (cudaStream_t)stream_);
}

int psb_trainer_result(psb_trainer* t, float* out3, int* num_rendered, void* stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (!t || t->geom_P < 0) { set_error_msg("psb_trainer_result: no step has run"); return PSB_ERR_ARG; }
	PSB_CUDA_OK(cudaEventSynchronize(t->readback));
	const uint32_t n = t->geom_P > 0 ? *t->h_count : 0;
	if (num_rendered) *num_rendered = (int)n;
	if (n > t->capacity) {
		// grow the binning arena; the step that just ran did not touch the parameters (its remaining kernels are no-ops,
		// but they may still be in flight: drain the stream before the arena is replaced)
		PSB_CUDA_OK(cudaStreamSynchronize(stream));
		PSB_CUDA_OK(cudaDeviceSynchronize());
		t->capacity = (size_t)(n * 1.25) + (1u << 16);
		t->bin_bytes = required_bytes<BinState>(t->capacity);
		int rc;
		if ((rc = dev_alloc(&t->bin_chunk, t->bin_bytes))) return rc;
		set_error_msg("psb_trainer: binning arena was too small for this view; it has been grown, repeat the step");
		return PSB_ERR_RETRY;
	}
	if (out3 && t->have_loss) {
		const double inv_n = 1.0 / (3.0 * (double)t->last_W * (double)t->last_H);
		const float l1 = (float)(t->h_sums[0] * inv_n), ss = (float)(t->h_sums[1] * inv_n);
		out3[0] = (1.0f - t->last_lambda) * l1 + t->last_lambda * (1.0f - ss);
		out3[1] = l1;
		out3[2] = ss;
	}
	return 0;
}

int psb_trainer_set_profiling(psb_trainer* t, int enable)
{
	if (!t) return PSB_ERR_ARG;
	if (enable && !t->ev_ready) {
		for (int i = 0; i <= psb_trainer::NSTAGE; i++) PSB_CUDA_OK(cudaEventCreate(&t->ev[i]));
		t->ev_ready = true;
	}
	t->profiling = enable != 0;
	t->ev_recorded = false;
	return 0;
}

// ms[0..6]: preprocess, depth sort + scan, binning (emit + tile sort + ranges), render forward, loss fwd+bwd,
// render backward, fused per-Gaussian backward + Adam — of the last profiled psb_trainer_step. Synchronises.
int psb_trainer_stage_times(psb_trainer* t, float* ms, int n)
{
	if (!t || !ms || n < 7 || !t->ev_recorded) { set_error_msg("psb_trainer_stage_times: no profiled step"); return PSB_ERR_ARG; }
	PSB_CUDA_OK(cudaEventSynchronize(t->ev[7]));
	for (int i = 0; i < 7; i++) PSB_CUDA_OK(cudaEventElapsedTime(&ms[i], t->ev[i], t->ev[i + 1]));
	return 0;
}

int psb_loss(int height, int width, const float* image, const float* gt_image, const float* mask, float lambda_dssim, float* dL_dimage,
             float* out3_host, void* stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (height <= 0 || width <= 0 || !image || !gt_image) { set_error_msg("psb_loss: bad argument"); return PSB_ERR_ARG; }
	float* dmap = nullptr;
	double* sums = nullptr;
	const size_t N = (size_t)height * width;
	PSB_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&dmap), 9 * N * sizeof(float)));
	PSB_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&sums), 2 * sizeof(double)));
	int rc = launch_loss(height, width, image, gt_image, mask, lambda_dssim, dmap, sums, dL_dimage, stream);
	double h[2] = {0, 0};
	if (rc == 0) {
		cudaMemcpyAsync(h, sums, sizeof(h), cudaMemcpyDeviceToHost, stream);
		if (cudaStreamSynchronize(stream) != cudaSuccess) rc = PSB_ERR_CUDA;
	}
	cudaFree(dmap); cudaFree(sums);
	if (rc == 0 && out3_host) {
		const double inv_n = 1.0 / (3.0 * (double)N);
		const float l1 = (float)(h[0] * inv_n), ss = (float)(h[1] * inv_n);
		out3_host[0] = (1.0f - lambda_dssim) * l1 + lambda_dssim * (1.0f - ss);
		out3_host[1] = l1;
		out3_host[2] = ss;
	}
	return rc;
}

}  // extern "C"
