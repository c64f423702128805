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
	uint32_t* h_count = nullptr; // pinned [4]: instance count of the last call | sticky overflow record {count, first seq, largest count}
	uint32_t* ovf = nullptr;     // device [4]: the sticky overflow record (tile_ranges_kernel), collected by psb_trainer_result
	uint32_t seq = 0;            // sequence number of the last forward pass (render / step / backward) of this context
	uint32_t ovf_first = 0, ovf_count = 0;  // what the last PSB_ERR_RETRY was about (psb_trainer_overflow_info)
	cudaEvent_t readback = nullptr;  // recorded once the instance count (and the loss sums) of the last call are in pinned memory
	int last_P = 0, last_W = 0, last_H = 0;
	float last_lambda = 0.2f;
	bool have_loss = false;
	// tight instance lists: only the tiles a splat can reach get an instance (see preprocess_fwd_kernel); PSB_TIGHT=0 keeps the
	// reference's full rectangles (A/B measurements only — images and gradients are the same either way)
	bool tight = !(getenv("PSB_TIGHT") && atoi(getenv("PSB_TIGHT")) == 0);
	// optional per-stage timing (CUDA events on the step's stream)
	bool profiling = false;
	// stage boundaries: 0 start | 1 preprocess | 2 depth sort | 3 binning | 4 render fwd | 5 loss | 6 render bwd |
	//   fused step:        7 per-Gaussian backward kernel | 8 f_rest Adam stream kernel
	//   data-parallel step: 7 push backward | 8 wait for every rank's records | 9 owner-side Adam (+ stores to every replica);
	//                       ev[NSTAGE] is recorded BEFORE the wait for the previous step's rows (stage "wait_params" = ev[NSTAGE] -> ev[0])
	static constexpr int NSTAGE = 11;   // (ev[10]: between the two owner-side Adam kernels of the data-parallel step)
	int last_stage = 0;  // index of the last boundary the last profiled call recorded
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
		PSB_CUDA_OK(cudaMallocHost(reinterpret_cast<void**>(&t->h_count), 4 * sizeof(uint32_t)));
		memset(t->h_count, 0, 4 * sizeof(uint32_t));
		if ((rc = dev_alloc(&t->ovf, 4))) return rc;
		const uint32_t init[4] = {0u, 0xFFFFFFFFu, 0u, 0u};
		PSB_CUDA_OK(cudaMemcpy(t->ovf, init, sizeof(init), cudaMemcpyHostToDevice));
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
	if ((rc = launch_scan_binning(P, cam, geom, bin, img, t->capacity, t->tight, t->ovf, ++t->seq, stream))) return rc;
	t->mark(3, stream);
	// the instance count travels to pinned memory as soon as it exists: psb_trainer_result waits on `readback`, not on the stream
	PSB_CUDA_OK(cudaMemcpyAsync(t->h_count, geom.counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
	PSB_CUDA_OK(cudaMemcpyAsync(t->h_count + 1, t->ovf, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
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
	rc = launch_fused_backward(grads == nullptr, 0, P, tt, cam, geom, t->sink, t->seeds, h, gs, st, geom.counters, (uint32_t)t->capacity, stream,
	                           (t->profiling && t->ev_ready) ? t->ev[7] : nullptr);
	t->mark(8, stream);
	t->last_stage = 8;
	t->ev_recorded = t->profiling && t->ev_ready;
	return rc;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------------
// NVLink data-parallel context (psb_dp_*): the symmetric arena of one rank and its peer mappings.
// ------------------------------------------------------------------------------------------------------------------------
struct psb_dp {
	int rank = 0, world = 1, P = 0;
	int nchunks = 0, nlocal_max = 0, nlocal = 0;
	char* base = nullptr;                       // local arena (cudaMalloc: exportable through CUDA IPC)
	char* peer[DP_MAX_WORLD] = {};              // arena of every rank in this process' address space (peer[rank] == base)
	size_t off_param[6] = {}, off_inbox = 0, off_meta = 0, off_gflag = 0, off_pflag = 0, bytes = 0;
	float* g_rest = nullptr;                    // local scratch
	uint32_t* local = nullptr;                  // [8]: 0 = backward done-counter, 1 = adam done-counter, 2 = status (1 = a wait timed out)
	uint32_t epoch = 0;
	bool connected = false;
	cudaIpcMemHandle_t handle;
	// pipeline: the Gaussians are cut into `groups` contiguous ranges (boundaries at multiples of 128 * world chunks); the owner-side Adam
	// of group g runs on `adam_stream` as soon as every rank's records of group g have landed, under the push backward of group g + 1
	int groups = 1, chunks_per_group = 0;
	cudaStream_t adam_stream = nullptr;
};

namespace {
constexpr size_t DP_ROW[6] = {3, 3, 45, 1, 3, 4};
void dp_layout(psb_dp* d)
{
	size_t o = 0;
	for (int i = 0; i < 6; i++) { d->off_param[i] = o; o = align_up(o + (size_t)d->P * DP_ROW[i] * sizeof(float), 256); }
	d->off_inbox = o; o = align_up(o + (size_t)d->world * d->nlocal_max * 128 * DP_REC * sizeof(float), 256);
	d->off_meta = o; o += 256;
	d->off_gflag = o; o += 256;
	d->off_pflag = o; o += 256;
	d->bytes = o;
}
}  // namespace

extern "C" {

int psb_dp_create(psb_dp** out, int rank, int world, int P)
{
	if (!out || world < 1 || world > DP_MAX_WORLD || rank < 0 || rank >= world || P < 0) { set_error_msg("psb_dp_create: need 0 <= rank < world <= 8, P >= 0"); return PSB_ERR_ARG; }
	psb_dp* d = new psb_dp();
	d->rank = rank; d->world = world; d->P = P;
	d->nchunks = (P + 127) / 128;
	d->nlocal_max = (d->nchunks + world - 1) / world;
	d->nlocal = d->nchunks > rank ? (d->nchunks - rank + world - 1) / world : 0;
	dp_layout(d);
	{
		const char* g = getenv("PSB_DP_GROUPS");
		// default 1 (no pipelining): measured on 2 and 8 B200s, running the owner-side Adam of group g under the push backward of group
		// g + 1 is slower than running them back to back (N=8: 3.32 ms/step with 1 group, 3.37 with 2, 3.46 with 4; N=2: 2.63 / 2.65 / 2.76 —
		// the two phases compete for the same store path, and the smaller launches lose their tails). Kept selectable for other fabrics.
		int G = g ? atoi(g) : 1;
		if (getenv("PSB_DP_SIGNAL") && strcmp(getenv("PSB_DP_SIGNAL"), "fence") == 0) G = 1;   // the in-kernel signalling variant is not pipelined
		G = G < 1 ? 1 : (G > 8 ? 8 : G);
		int cpg = (d->nchunks + G - 1) / G;
		cpg = (cpg + world - 1) / world * world;          // a group holds whole rounds of the chunk -> owner rotation
		if (cpg < world) cpg = world;
		d->chunks_per_group = cpg;
		d->groups = d->nchunks > 0 ? (d->nchunks + cpg - 1) / cpg : 1;
	}
	cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&d->base), d->bytes);
	if (e == cudaSuccess) e = cudaMemset(d->base, 0, d->bytes);
	if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d->g_rest), ((size_t)d->nlocal_max * 128 * 45 + 64) * sizeof(float));
	if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d->local), 8 * sizeof(uint32_t));
	if (e == cudaSuccess) e = cudaMemset(d->local, 0, 8 * sizeof(uint32_t));
	if (e == cudaSuccess && world > 1) e = cudaIpcGetMemHandle(&d->handle, d->base);
	if (e == cudaSuccess && d->groups > 1) e = cudaStreamCreateWithFlags(&d->adam_stream, cudaStreamNonBlocking);
	if (e != cudaSuccess) {
		set_error("psb_dp_create", e, __FILE__, __LINE__);
		cudaFree(d->base); cudaFree(d->g_rest); cudaFree(d->local);
		delete d;
		return PSB_ERR_CUDA;
	}
	d->peer[rank] = d->base;
	d->connected = world == 1;
	*out = d;
	return 0;
}

int psb_dp_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int psb_dp_ipc_handle(psb_dp* d, void* out_handle)
{
	if (!d || !out_handle) { set_error_msg("psb_dp_ipc_handle: null"); return PSB_ERR_ARG; }
	memcpy(out_handle, &d->handle, sizeof(cudaIpcMemHandle_t));
	return 0;
}

int psb_dp_connect(psb_dp* d, const void* handles)
{
	if (!d || (d->world > 1 && !handles)) { set_error_msg("psb_dp_connect: null"); return PSB_ERR_ARG; }
	for (int j = 0; j < d->world; j++) {
		if (j == d->rank || d->peer[j]) continue;
		cudaIpcMemHandle_t h;
		memcpy(&h, static_cast<const char*>(handles) + (size_t)j * sizeof(h), sizeof(h));
		void* p = nullptr;
		PSB_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
		d->peer[j] = static_cast<char*>(p);
	}
	d->connected = true;
	return 0;
}

int psb_dp_params(psb_dp* d, float** out6)
{
	if (!d || !out6) { set_error_msg("psb_dp_params: null"); return PSB_ERR_ARG; }
	for (int i = 0; i < 6; i++) out6[i] = reinterpret_cast<float*>(d->base + d->off_param[i]);
	return 0;
}

int psb_dp_destroy(psb_dp* d)
{
	if (!d) return 0;
	cudaDeviceSynchronize();
	for (int j = 0; j < d->world; j++)
		if (j != d->rank && d->peer[j]) cudaIpcCloseMemHandle(d->peer[j]);
	cudaFree(d->base); cudaFree(d->g_rest); cudaFree(d->local);
	if (d->adam_stream) cudaStreamDestroy(d->adam_stream);
	delete d;
	return 0;
}

// Enqueues a wait (device side) until every rank's parameter rows of the last psb_dp_step have landed in THIS rank's
// tensors: after it, the local replica is complete and may be read by anything that follows on `stream`.
int psb_dp_sync(psb_dp* d, void* stream_)
{
	if (!d) { set_error_msg("psb_dp_sync: null"); return PSB_ERR_ARG; }
	if (d->epoch == 0 || d->P == 0 || (d->world == 1 && d->groups == 1)) return 0;   // (one rank, one group: everything is on the caller's stream)
	return launch_wait_flags(reinterpret_cast<const uint32_t*>(d->base + d->off_pflag), d->groups * d->world, d->epoch, d->local + 2, (cudaStream_t)stream_);
}

// 0 = healthy, 1 = a cross-rank wait timed out (a peer never arrived); synchronises the stream
int psb_dp_status(psb_dp* d, void* stream_)
{
	if (!d) { set_error_msg("psb_dp_status: null"); return PSB_ERR_ARG; }
	uint32_t st = 0;
	PSB_CUDA_OK(cudaMemcpyAsync(&st, d->local + 2, sizeof(st), cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
	PSB_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream_));
	return (int)st;
}

int psb_dp_step(psb_trainer* t, psb_dp* d, int P, int M, const psb_model* model, const psb_camera* camera, const float* background,
                const float* gt_image, const float* mask, const psb_step* step, float* out_color, int* radii, void* stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	int rc;
	if (!d || !d->connected || P != d->P || !model) { set_error_msg("psb_dp_step: context not connected or built for a different P"); return PSB_ERR_ARG; }
	if ((rc = check_model(P, M, model, true))) return rc;
	for (int i = 0; i < 6; i++)
		if (P > 0 && model->param[i] != reinterpret_cast<float*>(d->base + d->off_param[i])) {
			set_error_msg("psb_dp_step: the model's parameter tensors must be the arena tensors returned by psb_dp_params");
			return PSB_ERR_ARG;
		}
	const uint32_t epoch = ++d->epoch;
	t->mark(psb_trainer::NSTAGE, stream);
	// every rank's updated rows of the previous step must have landed here before this step reads the parameters
	if (epoch > 1 && P > 0 && (d->world > 1 || d->groups > 1))   // (also orders this step behind the previous step's work on the adam stream)
		if ((rc = launch_wait_flags(reinterpret_cast<const uint32_t*>(d->base + d->off_pflag), d->groups * d->world, epoch - 1, d->local + 2, stream))) return rc;
	// render, loss, tile backward (the 9 screen-space sums per Gaussian)
	if ((rc = step_impl(t, P, M, model, camera, background, gt_image, mask, step, out_color, radii, nullptr, stream, /*tiles_only=*/true))) return rc;
	if (P == 0) return 0;
	char* gc = t->geom_chunk;
	GeomState geom = GeomState::from_chunk(gc, (size_t)P);
	const Camera cam = to_camera(camera);
	TrainTensors tt;
	for (int i = 0; i < 6; i++) { tt.p[i] = model->param[i]; tt.m[i] = model->exp_avg[i]; tt.v[i] = model->exp_avg_sq[i]; }
	DensifyStats st;
	st.enabled = (step->update_densify_stats && model->max_radii2D && model->xyz_gradient_accum && model->denom) ? 1 : 0;
	st.max_radii2D = model->max_radii2D; st.xyz_gradient_accum = model->xyz_gradient_accum; st.denom = model->denom;
	const StepHyper h = to_hyper(step);
	DpPush push;
	memset(&push, 0, sizeof(push));
	DpShard sh;
	memset(&sh, 0, sizeof(sh));
	for (int j = 0; j < d->world; j++) {
		push.inbox[j] = reinterpret_cast<float*>(d->peer[j] + d->off_inbox);
		push.meta[j] = reinterpret_cast<float*>(d->peer[j] + d->off_meta);
		push.grad_flag[j] = reinterpret_cast<uint32_t*>(d->peer[j] + d->off_gflag);
		sh.param_flag[j] = reinterpret_cast<uint32_t*>(d->peer[j] + d->off_pflag);
		for (int i = 0; i < 6; i++) sh.param[j][i] = reinterpret_cast<float*>(d->peer[j] + d->off_param[i]);
	}
	push.done_counter = d->local; push.world = d->world; push.rank = d->rank; push.nlocal_max = d->nlocal_max; push.epoch = epoch;
	static const int fence_in_kernel = (getenv("PSB_DP_SIGNAL") && strcmp(getenv("PSB_DP_SIGNAL"), "fence") == 0) ? 1 : 0;
	push.fence_in_kernel = fence_in_kernel; sh.fence_in_kernel = fence_in_kernel;
	sh.inbox = reinterpret_cast<const float*>(d->base + d->off_inbox);
	sh.meta = reinterpret_cast<const float*>(d->base + d->off_meta);
	sh.g_rest = d->g_rest; sh.done_counter = d->local + 1;
	sh.world = d->world; sh.rank = d->rank; sh.nlocal_max = d->nlocal_max; sh.P = P; sh.epoch = epoch;
	static const int rotate = (getenv("PSB_DP_ROTATE") && atoi(getenv("PSB_DP_ROTATE")) == 0) ? 0 : 1;
	sh.rotate = rotate;
	// TMA bulk stores for the small-parameter chunks: verified over 2 GPUs (same result, same time as per-thread stores); not yet measured
	// on 8 — the default stays the configuration the 8-GPU numbers in profiles/ were taken with
	static const int bulk = (getenv("PSB_DP_BULK") && atoi(getenv("PSB_DP_BULK")) != 0) ? 1 : 0;
	sh.bulk = bulk;
	// Pipeline over groups of chunks: [main stream] push backward of group g (records go straight into the owners' inboxes) + signal;
	// [adam stream] wait until every rank's records of group g have landed -> Adam of the owned rows of group g, updated rows stored to
	// every replica + signal. With one group everything stays on the main stream. Flag word of (group g, rank r) = g * world + r.
	cudaStream_t as = d->groups > 1 ? d->adam_stream : stream;
	const int rows_per_group = d->chunks_per_group * 128;
	for (int g = 0; g < d->groups; g++) {
		DpPush pg = push;
		for (int j = 0; j < d->world; j++) pg.grad_flag[j] = push.grad_flag[j] + g * d->world;
		const int first = g * rows_per_group, last = (g + 1) * rows_per_group < P ? (g + 1) * rows_per_group : P;
		if ((rc = launch_push_backward(first, last, tt, cam, geom, t->sink, h, st, geom.counters, (uint32_t)t->capacity, pg, stream))) return rc;
	}
	t->mark(7, stream);
	t->mark(8, stream);
	for (int g = 0; g < d->groups; g++) {
		if ((rc = launch_wait_flags(reinterpret_cast<const uint32_t*>(d->base + d->off_gflag) + g * d->world, d->world, epoch, d->local + 2, as))) return rc;
		DpShard sg = sh;
		for (int j = 0; j < d->world; j++) sg.param_flag[j] = sh.param_flag[j] + g * d->world;
		sg.lc_first = g * (d->chunks_per_group / d->world);
		const int lc_end = (g + 1) * (d->chunks_per_group / d->world) < d->nlocal ? (g + 1) * (d->chunks_per_group / d->world) : d->nlocal;
		sg.nlocal = lc_end > sg.lc_first ? lc_end - sg.lc_first : 0;
		if ((rc = launch_shard_adam(sg, tt, h, 1.0f / (float)d->world, as, (t->profiling && t->ev_ready && g == d->groups - 1) ? t->ev[10] : nullptr))) return rc;
	}
	t->mark(9, as);
	t->last_stage = 9;
	t->ev_recorded = t->profiling && t->ev_ready;
	return rc;
}

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
	cudaFree(t->sink); cudaFree(t->seeds); cudaFree(t->sums); cudaFree(t->ovf);
	if (t->ev_ready) for (int i = 0; i <= psb_trainer::NSTAGE; i++) cudaEventDestroy(t->ev[i]);
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
	if (P > 0 && ((reinterpret_cast<uintptr_t>(grads[2]) | reinterpret_cast<uintptr_t>(grads[5])) & 15)) {
		set_error_msg("psb_trainer_backward: the features_rest and rotation gradient segments must be 16-byte aligned");
		return PSB_ERR_ARG;
	}
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
	if ((reinterpret_cast<uintptr_t>(grads[2]) | reinterpret_cast<uintptr_t>(grads[5])) & 15) {
		set_error_msg("psb_trainer_backward_slab: the features_rest and rotation gradient blocks must be 16-byte aligned");
		return PSB_ERR_ARG;
	}
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
	const uint32_t n = t->geom_P > 0 ? t->h_count[0] : 0;
	if (num_rendered) *num_rendered = (int)n;
	const uint32_t n_ovf = t->geom_P > 0 ? t->h_count[1] : 0;
	if (n_ovf > 0) {
		// One or more views since the last collection needed more instances than the binning arena holds: those calls were
		// no-ops on the model (fused step) / contributed zero gradients (data-parallel modes). The record is sticky on the device,
		// so a queued overflow is reported here even when later steps fitted. Grow the arena (the remaining kernels may still be
		// in flight: drain first), clear the record, tell the caller which call it was.
		PSB_CUDA_OK(cudaStreamSynchronize(stream));
		PSB_CUDA_OK(cudaDeviceSynchronize());
		uint32_t rec[3];
		PSB_CUDA_OK(cudaMemcpy(rec, t->ovf, sizeof(rec), cudaMemcpyDeviceToHost));
		t->ovf_count = rec[0]; t->ovf_first = rec[1];
		const uint32_t init[4] = {0u, 0xFFFFFFFFu, 0u, 0u};
		PSB_CUDA_OK(cudaMemcpy(t->ovf, init, sizeof(init), cudaMemcpyHostToDevice));
		t->h_count[1] = 0;
		t->capacity = (size_t)(rec[2] * 1.25) + (1u << 16);
		t->bin_bytes = required_bytes<BinState>(t->capacity);
		int rc;
		if ((rc = dev_alloc(&t->bin_chunk, t->bin_bytes))) return rc;
		set_error_msg("psb_trainer: binning arena was too small for a view; it has been grown, repeat the call (psb_trainer_overflow_info says which)");
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

// Parity/debug export (tests only): per pixel, the final transmittance and the Gaussian index of the last blended splat
// (-1: none) of the last step / render. List positions (n_contrib) are private — tight lists number the instances
// differently from the reference — but the splat they name is comparable with the reference's
// point_list[ranges[tile].x + n_contrib - 1] (cuda_rasterizer/forward.cu:352-365).
__global__ void export_last_splat_kernel(int W, int H, int grid_x, const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                                         const uint32_t* __restrict__ n_contrib, const float* __restrict__ final_T, int* __restrict__ last_gauss,
                                         float* __restrict__ T_out)
{
	const int pix = blockIdx.x * blockDim.x + threadIdx.x;
	if (pix >= W * H) return;
	const int x = pix % W, y = pix / W;
	const uint2 r = ranges[(y / PSB_TILE_Y) * grid_x + x / PSB_TILE_X];
	const uint32_t n = n_contrib[pix];
	if (last_gauss) last_gauss[pix] = n ? (int)point_list[r.x + n - 1] : -1;
	if (T_out) T_out[pix] = final_T[pix];
}

int psb_trainer_debug_state(psb_trainer* t, int width, int height, int* last_gauss, float* final_T, int* num_rendered, void* stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (!t || t->geom_P < 0 || t->last_W != width || t->last_H != height) { set_error_msg("psb_trainer_debug_state: no step of that size has run"); return PSB_ERR_ARG; }
	char* gc = t->geom_chunk; GeomState geom = GeomState::from_chunk(gc, (size_t)t->geom_P);
	char* ic = t->img_chunk; ImgState img = ImgState::from_chunk(ic, (size_t)width * height);
	char* bc = t->bin_chunk; BinState bin = BinState::from_chunk(bc, t->capacity);
	const int grid_x = (width + PSB_TILE_X - 1) / PSB_TILE_X, grid_y = (height + PSB_TILE_Y - 1) / PSB_TILE_Y;
	const int res = make_sort_plan(tile_id_bits(grid_x * grid_y)).npass & 1;
	export_last_splat_kernel<<<(width * height + 255) / 256, 256, 0, stream>>>(width, height, grid_x, img.ranges, bin.inst[res], img.n_contrib, img.final_T,
	                                                                        last_gauss, final_T);
	PSB_LAUNCH_OK();
	if (num_rendered) {
		PSB_CUDA_OK(cudaMemcpyAsync(num_rendered, geom.counters, sizeof(int), cudaMemcpyDeviceToHost, stream));
		PSB_CUDA_OK(cudaStreamSynchronize(stream));
	}
	return 0;
}

int psb_trainer_overflow_info(psb_trainer* t, unsigned* first_seq, unsigned* count, unsigned* current_seq)
{
	if (!t) { set_error_msg("psb_trainer_overflow_info: null"); return PSB_ERR_ARG; }
	if (first_seq) *first_seq = t->ovf_first;
	if (count) *count = t->ovf_count;
	if (current_seq) *current_seq = t->seq;
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

// ms[i] = time between stage boundaries i and i+1 of the last profiled step (see psb_trainer::ev), i < n; entries beyond the last
// recorded boundary are 0. Fused step: 8 stages (preprocess, depth sort, binning, render fwd, loss, render bwd, per-Gaussian
// backward kernel, f_rest Adam kernel). Data-parallel step: 9 stages (..., render bwd, push backward, wait for the ranks' records,
// owner-side Adam) and ms[9] = the wait for the previous step's rows at the start of the step. Synchronises. Returns the number of
// stages filled.
int psb_trainer_stage_times(psb_trainer* t, float* ms, int n)
{
	if (!t || !ms || n < 7 || !t->ev_recorded) { set_error_msg("psb_trainer_stage_times: no profiled step"); return PSB_ERR_ARG; }
	PSB_CUDA_OK(cudaEventSynchronize(t->ev[t->last_stage]));
	for (int i = 0; i < n; i++) ms[i] = 0.f;
	int filled = 0;
	for (int i = 0; i < t->last_stage && i < n; i++, filled++) PSB_CUDA_OK(cudaEventElapsedTime(&ms[i], t->ev[i], t->ev[i + 1]));
	if (t->last_stage == 9 && n > 9) { PSB_CUDA_OK(cudaEventElapsedTime(&ms[9], t->ev[psb_trainer::NSTAGE], t->ev[0])); filled = 10; }
	if (t->last_stage == 9 && n > 10) { PSB_CUDA_OK(cudaEventElapsedTime(&ms[10], t->ev[10], t->ev[9])); filled = 11; }   // the f_rest kernel + signal alone
	return filled;
}

int psb_loss(int height, int width, const float* image, const float* gt_image, const float* mask, float lambda_dssim, float* dL_dimage,
             float* out3_host, void* stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (height <= 0 || width <= 0 || !image || !gt_image) { set_error_msg("psb_loss: bad argument"); return PSB_ERR_ARG; }
	float* dmap = nullptr;
	double* sums = nullptr;
	const size_t N = (size_t)height * width;
	// stream-ordered scratch: no device-wide synchronisation, freed on every path
	PSB_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&dmap), 9 * N * sizeof(float), stream));
	if (cudaMallocAsync(reinterpret_cast<void**>(&sums), 2 * sizeof(double), stream) != cudaSuccess) {
		cudaFreeAsync(dmap, stream);
		set_error_msg("psb_loss: out of device memory");
		return PSB_ERR_CUDA;
	}
	int rc = launch_loss(height, width, image, gt_image, mask, lambda_dssim, dmap, sums, dL_dimage, stream);
	double h[2] = {0, 0};
	if (rc == 0) {
		cudaMemcpyAsync(h, sums, sizeof(h), cudaMemcpyDeviceToHost, stream);
		if (cudaStreamSynchronize(stream) != cudaSuccess) rc = PSB_ERR_CUDA;
	}
	cudaFreeAsync(dmap, stream); cudaFreeAsync(sums, stream);
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
