"""Host-side mirror of the reference's trainer-level surface for the hot path, on the psb200 C-ABI.

  GaussianOptimizationParams   reference include/gaussian_parameters.h:49-96 (values: cfg/gaussian_mapper/RGB-D/Replica/replica_rgbd.yaml:53-74)
  GaussianModel                reference include/gaussian_model.h:59-193 — the tensors, Adam state, learning-rate
                               handling (trainingSetup / updateLearningRate / set*LearningRate / exponLrFunc,
                               src/gaussian_model.cpp:477-554, 1118-1131)
  GaussianTrainer              reference GaussianMapper::trainForOneIteration (src/gaussian_mapper.cpp:614-774) /
                               GaussianTrainer::trainingOnce (src/gaussian_trainer.cpp:31-135): one iteration =
                               render, L1 + DSSIM, backward, densification statistics, Adam step — here ONE call
                               into psb_trainer_step.
PyTorch provides device memory, streams and (for the data-parallel variant) torch.distributed.
"""
import ctypes as C
import math
from dataclasses import dataclass

import torch

from . import _lib

GROUPS = ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation")


class _Model(C.Structure):
    _fields_ = [("param", C.c_void_p * 6), ("exp_avg", C.c_void_p * 6), ("exp_avg_sq", C.c_void_p * 6),
                ("max_radii2D", C.c_void_p), ("xyz_gradient_accum", C.c_void_p), ("denom", C.c_void_p), ("exist_since_iter", C.c_void_p)]


class _Camera(C.Structure):
    _fields_ = [("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p), ("tan_fovx", C.c_float),
                ("tan_fovy", C.c_float), ("width", C.c_int), ("height", C.c_int)]


class _Step(C.Structure):
    _fields_ = [("lr", C.c_float * 6), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("step", C.c_int),
                ("lambda_dssim", C.c_float), ("sh_degree", C.c_int), ("update_densify_stats", C.c_int)]


class _DensifyCfg(C.Structure):
    _fields_ = [("max_grad", C.c_float), ("min_opacity", C.c_float), ("extent", C.c_float), ("percent_dense", C.c_float), ("max_screen_size", C.c_int),
                ("seed", C.c_ulonglong), ("offset", C.c_ulonglong)]


def _bind():
    L = _lib.lib()
    if getattr(L, "_trainer_bound", False):
        return L
    vp = C.c_void_p
    L.psb_trainer_create.argtypes = [C.POINTER(vp)]
    L.psb_trainer_destroy.argtypes = [vp]
    L.psb_trainer_render.argtypes = [vp, C.c_int, C.c_int, C.POINTER(_Model), C.POINTER(_Camera), vp, C.c_int, vp, vp, vp]
    L.psb_trainer_step.argtypes = [vp, C.c_int, C.c_int, C.POINTER(_Model), C.POINTER(_Camera), vp, vp, vp, C.POINTER(_Step), vp, vp, vp]
    L.psb_trainer_backward.argtypes = [vp, C.c_int, C.c_int, C.POINTER(_Model), C.POINTER(_Camera), vp, vp, vp, C.POINTER(_Step), vp, vp,
                                       C.POINTER(vp), vp]
    L.psb_adam_update.argtypes = [C.c_int, C.c_int, C.POINTER(_Model), C.POINTER(vp), C.POINTER(_Step), C.c_float, vp]
    L.psb_trainer_result.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int), vp]
    L.psb_loss.argtypes = [C.c_int, C.c_int, vp, vp, vp, C.c_float, vp, C.POINTER(C.c_float), vp]
    for n in ("psb_trainer_create", "psb_trainer_destroy", "psb_trainer_render", "psb_trainer_step", "psb_trainer_backward",
              "psb_adam_update", "psb_trainer_result", "psb_loss"):
        getattr(L, n).restype = C.c_int
    L.psb_densify_workspace_bytes.argtypes = [C.c_int]
    L.psb_densify_workspace_bytes.restype = C.c_size_t
    L.psb_densify_plan.argtypes = [C.c_int, C.POINTER(_Model), C.POINTER(_DensifyCfg), vp, C.POINTER(C.c_int), vp]
    L.psb_densify_apply.argtypes = [C.c_int, C.POINTER(_Model), C.POINTER(_Model), C.c_int, C.POINTER(_DensifyCfg), vp, vp, vp]
    L.psb_prune_plan.argtypes = [C.c_int, vp, vp, C.POINTER(C.c_int), vp]
    L.psb_prune_apply.argtypes = [C.c_int, C.POINTER(_Model), C.POINTER(_Model), C.c_int, vp, C.c_int, vp]
    L.psb_insert_points.argtypes = [C.c_int, C.POINTER(_Model), C.POINTER(_Model), C.c_int, vp, vp, vp, C.c_int, vp]
    L.psb_reset_opacity.argtypes = [C.c_int, vp, vp, vp, vp]
    for n in ("psb_densify_plan", "psb_densify_apply", "psb_prune_plan", "psb_prune_apply", "psb_insert_points", "psb_reset_opacity"):
        getattr(L, n).restype = C.c_int
    L._trainer_bound = True
    return L


@dataclass
class GaussianOptimizationParams:
    iterations: int = 30100
    position_lr_init: float = 0.00032
    position_lr_final: float = 0.00032
    position_lr_delay_mult: float = 0.01
    position_lr_max_steps: int = 24
    feature_lr: float = 0.0025
    opacity_lr: float = 0.05
    scaling_lr: float = 0.005
    rotation_lr: float = 0.001
    percent_dense: float = 0.01
    lambda_dssim: float = 0.2
    densification_interval: int = 100
    opacity_reset_interval: int = 0
    densify_from_iter: int = 600
    densify_until_iter: int = 15000
    densify_grad_threshold: float = 0.001


class GaussianModel:
    """Parameter tensors in the reference's layout + Adam state + learning rates."""

    def __init__(self, sh_degree=3, device="cuda"):
        self.max_sh_degree_ = sh_degree
        self.active_sh_degree_ = 0
        self.device = torch.device(device)
        self.spatial_lr_scale_ = 1.0
        self.lr_ = [0.0] * 6
        self.step_ = 0
        self.lr_delay_steps_ = 0
        self.percent_dense_ = 0.01

    @classmethod
    def from_numpy(cls, scene, device="cuda", sh_degree=3):
        m = cls(sh_degree, device)
        t = lambda a: torch.from_numpy(a).to(m.device).contiguous()
        m.xyz_, m.features_dc_, m.features_rest_ = t(scene["xyz"]), t(scene["features_dc"]), t(scene["features_rest"])
        m.opacity_, m.scaling_, m.rotation_ = t(scene["opacity"]), t(scene["scaling"]), t(scene["rotation"])
        m.active_sh_degree_ = sh_degree
        return m

    def tensors(self):
        return [self.xyz_, self.features_dc_, self.features_rest_, self.opacity_, self.scaling_, self.rotation_]

    def num_points(self):
        return self.xyz_.size(0)

    # --- reference GaussianModel::trainingSetup (gaussian_model.cpp:477-509)
    def trainingSetup(self, args: GaussianOptimizationParams):
        self.percent_dense_ = args.percent_dense
        P = self.num_points()
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=self.device)
        self.xyz_gradient_accum_, self.denom_, self.max_radii2D_ = z(P, 1), z(P, 1), z(P)
        self.exp_avg_ = [torch.zeros_like(t) for t in self.tensors()]
        self.exp_avg_sq_ = [torch.zeros_like(t) for t in self.tensors()]
        self.step_ = 0
        self.lr_ = [args.position_lr_init * self.spatial_lr_scale_, args.feature_lr, args.feature_lr / 20.0, args.opacity_lr,
                    args.scaling_lr, args.rotation_lr]
        self.lr_init_ = args.position_lr_init * self.spatial_lr_scale_
        self.lr_final_ = args.position_lr_final * self.spatial_lr_scale_
        self.lr_delay_mult_ = args.position_lr_delay_mult
        self.max_steps_ = args.position_lr_max_steps

    # --- training-state snapshot (parameters, Adam moments, densification statistics, step counter); restore copies in
    #     place so device pointers held elsewhere stay valid. Used by bench.py to time every leg on the same iterations.
    def snapshot(self):
        c = lambda ts: [t.clone() for t in ts]
        return dict(p=c(self.tensors()), m=c(self.exp_avg_), v=c(self.exp_avg_sq_),
                    stats=c([self.xyz_gradient_accum_, self.denom_, self.max_radii2D_]), step=self.step_, lr=list(self.lr_),
                    sh=self.active_sh_degree_)

    def restore(self, snap):
        assert snap["p"][0].shape == self.xyz_.shape, "restore: the model was resized since the snapshot"
        for dst, src in zip(self.tensors() + self.exp_avg_ + self.exp_avg_sq_ + [self.xyz_gradient_accum_, self.denom_, self.max_radii2D_],
                            snap["p"] + snap["m"] + snap["v"] + snap["stats"]):
            dst.copy_(src)
        self.step_, self.lr_, self.active_sh_degree_ = snap["step"], list(snap["lr"]), snap["sh"]

    # --- on-disk point cloud (reference GaussianModel::savePly / loadPly / saveSparsePointsPly, gaussian_model.cpp:838-1088)
    def savePly(self, path):
        from . import io
        io.save_ply(path, *[t.detach().cpu().numpy() for t in self.tensors()])

    def loadPly(self, path):
        """Replaces the parameter tensors by the file's content and activates every SH degree (gaussian_model.cpp:951)."""
        from . import io
        a = io.load_ply(path, self.max_sh_degree_)
        t = lambda x: torch.from_numpy(x).to(self.device).contiguous()
        self._set([t(a[k]) for k in ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation")])
        self.active_sh_degree_ = self.max_sh_degree_

    def setSparsePoints(self, xyz, color):
        """sparse SLAM map points + colours in [0,1] kept for input.ply (reference gaussian_model.cpp: sparse_points_xyz_ / _color_)"""
        self.sparse_points_xyz_, self.sparse_points_color_ = xyz, color

    def saveSparsePointsPly(self, path):
        from . import io
        io.save_sparse_points_ply(path, self.sparse_points_xyz_.cpu().numpy(), self.sparse_points_color_.cpu().numpy())

    def exponLrFunc(self, step):
        if step < 0 or (self.lr_init_ == 0.0 and self.lr_final_ == 0.0):
            return 0.0
        if self.lr_delay_steps_ > 0:
            delay = self.lr_delay_mult_ + (1.0 - self.lr_delay_mult_) * math.sin(0.5 * math.pi * min(max(step / self.lr_delay_steps_, 0.0), 1.0))
        else:
            delay = 1.0
        t = min(max(step / self.max_steps_, 0.0), 1.0)
        return delay * math.exp(math.log(self.lr_init_) * (1 - t) + math.log(self.lr_final_) * t)

    def updateLearningRate(self, step):
        self.lr_[0] = self.exponLrFunc(step)
        return self.lr_[0]

    def setPositionLearningRate(self, lr): self.lr_[0] = lr * self.spatial_lr_scale_
    def setFeatureLearningRate(self, lr): self.lr_[1], self.lr_[2] = lr, lr / 20.0
    def setOpacityLearningRate(self, lr): self.lr_[3] = lr
    def setScalingLearningRate(self, lr): self.lr_[4] = lr
    def setRotationLearningRate(self, lr): self.lr_[5] = lr
    def setShDegree(self, d): self.active_sh_degree_ = min(max(d, 0), self.max_sh_degree_)

    def oneUpShDegree(self):
        if self.active_sh_degree_ < self.max_sh_degree_:
            self.active_sh_degree_ += 1

    # ------------------------------------------------------------------------------------------------------------
    # Densification / pruning / insertion / loop-closure surgery (reference src/gaussian_model.cpp:193-475, 556-815).
    # Every one of them is a fused device pass over the model through the C-ABI (psb_densify_* / psb_prune_* /
    # psb_insert_points / psb_reset_opacity): one read of the surviving rows, one write of the output rows, one host
    # round trip (the new row count). The ATen restatement of the reference code lives in oracle/ref_densify.py (tests only).
    # ------------------------------------------------------------------------------------------------------------
    def getOpacityActivation(self): return torch.sigmoid(self.opacity_)
    def getScalingActivation(self): return torch.exp(self.scaling_)
    def getRotationActivation(self): return torch.nn.functional.normalize(self.rotation_)

    def _set(self, tensors):
        (self.xyz_, self.features_dc_, self.features_rest_, self.opacity_, self.scaling_, self.rotation_) = tensors

    def _exist(self):
        if getattr(self, "exist_since_iter_", None) is None or self.exist_since_iter_.size(0) != self.num_points():
            self.exist_since_iter_ = torch.zeros(self.num_points(), dtype=torch.int32, device=self.device)
        return self.exist_since_iter_

    def _blank(self, P):
        """Uninitialised tensors of a P-row model (parameters, both moments, statistics, exist_since_iter)."""
        M = (self.max_sh_degree_ + 1) ** 2
        shapes = [(P, 3), (P, 1, 3), (P, M - 1, 3), (P, 1), (P, 3), (P, 4)]
        e = lambda sh: torch.empty(sh, dtype=torch.float32, device=self.device)
        return dict(p=[e(sh) for sh in shapes], m=[e(sh) for sh in shapes], v=[e(sh) for sh in shapes], accum=e((P, 1)), denom=e((P, 1)),
                    max_radii=e((P,)), exist=torch.empty(P, dtype=torch.int32, device=self.device))

    @staticmethod
    def _cm(p, m, v, accum, denom, max_radii, exist):
        c = _Model()
        for i in range(6):
            c.param[i], c.exp_avg[i], c.exp_avg_sq[i] = p[i].data_ptr(), m[i].data_ptr(), v[i].data_ptr()
        c.max_radii2D, c.xyz_gradient_accum, c.denom, c.exist_since_iter = max_radii.data_ptr(), accum.data_ptr(), denom.data_ptr(), exist.data_ptr()
        return c

    def _src(self):
        self._exist()
        return self._cm(self.tensors(), self.exp_avg_, self.exp_avg_sq_, self.xyz_gradient_accum_, self.denom_, self.max_radii2D_, self.exist_since_iter_)

    def _adopt(self, d):
        self._set(d["p"])
        self.exp_avg_, self.exp_avg_sq_ = d["m"], d["v"]
        self.xyz_gradient_accum_, self.denom_, self.max_radii2D_, self.exist_since_iter_ = d["accum"], d["denom"], d["max_radii"], d["exist"]

    def resetOpacity(self):
        """reference gaussian_model.cpp:556-565. NOTE: the reference's misplaced parenthesis makes this
        min(sigmoid(o), 1) — it does NOT clamp to 0.01; only the opacity moments are zeroed."""
        L = _bind()
        _lib.check(L.psb_reset_opacity(self.num_points(), self.opacity_.data_ptr(), self.exp_avg_[3].data_ptr(), self.exp_avg_sq_[3].data_ptr(),
                                       torch.cuda.current_stream().cuda_stream), "psb_reset_opacity")

    def prunePoints(self, mask):
        """reference gaussian_model.cpp:588-642 (statistics are compacted, not reset)"""
        L = _bind()
        P = self.num_points()
        stream = torch.cuda.current_stream().cuda_stream
        ws = torch.empty(L.psb_densify_workspace_bytes(P), dtype=torch.uint8, device=self.device)
        counts = (C.c_int * 5)()
        m8 = mask.to(torch.uint8).contiguous()
        _lib.check(L.psb_prune_plan(P, m8.data_ptr(), ws.data_ptr(), counts, stream), "psb_prune_plan")
        d = self._blank(counts[0])
        src, dst = self._src(), self._cm(d["p"], d["m"], d["v"], d["accum"], d["denom"], d["max_radii"], d["exist"])
        _lib.check(L.psb_prune_apply(P, C.byref(src), C.byref(dst), counts[0], ws.data_ptr(), 1, stream), "psb_prune_apply")
        self._adopt(d)

    def densifyAndPrune(self, max_grad, min_opacity, extent, max_screen_size, seed=0, offset=None, samples=None):
        """reference gaussian_model.cpp:795-815 with densifyAndClone / densifyAndSplit / prunePoints / densificationPostfix folded
        into one plan + one scatter pass. The split draw is Philox(seed, offset) — identical on every data-parallel replica —
        unless `samples` ([2 * n_split, 3] standard normals in the reference's repeat order) is injected (tests).
        Returns (P_new, kept originals, clones, kept split children per copy, split-selected)."""
        L = _bind()
        P = self.num_points()
        stream = torch.cuda.current_stream().cuda_stream
        cfg = _DensifyCfg(float(max_grad), float(min_opacity), float(extent), float(self.percent_dense_), int(max_screen_size or 0), int(seed),
                          int(self.step_ if offset is None else offset))
        ws = torch.empty(L.psb_densify_workspace_bytes(P), dtype=torch.uint8, device=self.device)
        counts = (C.c_int * 5)()
        src = self._src()
        _lib.check(L.psb_densify_plan(P, C.byref(src), C.byref(cfg), ws.data_ptr(), counts, stream), "psb_densify_plan")
        if samples is not None:
            samples = samples.to(self.device, torch.float32).contiguous()
            assert samples.shape == (2 * counts[4], 3), (samples.shape, counts[4])
        d = self._blank(counts[0])
        dst = self._cm(d["p"], d["m"], d["v"], d["accum"], d["denom"], d["max_radii"], d["exist"])
        _lib.check(L.psb_densify_apply(P, C.byref(src), C.byref(dst), counts[0], C.byref(cfg), ws.data_ptr(),
                                       samples.data_ptr() if samples is not None and samples.numel() else None, stream), "psb_densify_apply")
        self._adopt(d)
        return tuple(counts)

    def densifySplitCount(self, max_grad, extent):
        """number of rows densifyAndSplit would select right now (plan only; for tests that inject the normal draw)"""
        L = _bind()
        P = self.num_points()
        cfg = _DensifyCfg(float(max_grad), 0.0, float(extent), float(self.percent_dense_), 0, 0, 0)
        ws = torch.empty(L.psb_densify_workspace_bytes(P), dtype=torch.uint8, device=self.device)
        counts = (C.c_int * 5)()
        src = self._src()
        _lib.check(L.psb_densify_plan(P, C.byref(src), C.byref(cfg), ws.data_ptr(), counts, torch.cuda.current_stream().cuda_stream), "psb_densify_plan")
        return counts[4]

    def increasePcd(self, points, colors, iteration=0):
        """reference gaussian_model.cpp:193-377 (both overloads: host vectors / tensors): new Gaussians from sparse points — RGB2SH
        colour, scale from the 3-NN mean distance (distCUDA2 = psb_dist_cuda2), identity rotation, opacity = logit(0.1)."""
        from .points import distCUDA2
        L = _bind()
        points = torch.as_tensor(points, dtype=torch.float32).reshape(-1, 3)
        colors = torch.as_tensor(colors, dtype=torch.float32).reshape(-1, 3)
        n = points.size(0)
        if n == 0:
            return
        pts = points.to(self.device).contiguous()
        cols = colors.to(self.device).contiguous()
        if getattr(self, "sparse_points_xyz_", None) is None or self.sparse_points_xyz_.numel() == 0:
            self.sparse_points_xyz_, self.sparse_points_color_ = pts, cols
        else:
            self.sparse_points_xyz_ = torch.cat((self.sparse_points_xyz_, pts), dim=0)
            self.sparse_points_color_ = torch.cat((self.sparse_points_color_, cols), dim=0)
        dist2 = distCUDA2(pts)
        P = self.num_points()
        if not hasattr(self, "exp_avg_"):
            raise RuntimeError("increasePcd: call trainingSetup() first (the optimizer state is extended together with the parameters)")
        d = self._blank(P + n)
        src, dst = self._src(), self._cm(d["p"], d["m"], d["v"], d["accum"], d["denom"], d["max_radii"], d["exist"])
        _lib.check(L.psb_insert_points(P, C.byref(src), C.byref(dst), n, pts.data_ptr(), cols.data_ptr(), dist2.data_ptr(), int(iteration),
                                       torch.cuda.current_stream().cuda_stream), "psb_insert_points")
        self._adopt(d)

    # --- loop closure / scale refinement (reference gaussian_model.cpp:379-475)
    def scaledTransformationPostfix(self, new_xyz, new_scaling):
        """:403-418: the two tensors replace xyz / scaling in the optimizer with FRESH (zero) moments (replaceTensorToOptimizer :567-586)"""
        self.xyz_, self.scaling_ = new_xyz.contiguous(), new_scaling.contiguous()
        for i in (0, 4):
            self.exp_avg_[i] = torch.zeros_like(self.tensors()[i])
            self.exp_avg_sq_[i] = torch.zeros_like(self.tensors()[i])

    def applyScaledTransformation(self, s, T):
        """:379-401: xyz <- T (s * xyz) through transformPoints; scaling_ (the LOG scale) is multiplied by s like the reference does.
        T: [4,4] row-major world transform (Sophus::SE3f::matrix())."""
        from .points import transformPoints
        T_tensor = torch.as_tensor(T, dtype=torch.float32, device=self.device).reshape(4, 4).transpose(0, 1).contiguous()
        xyz = transformPoints(self.xyz_ * float(s), T_tensor)
        self.scaledTransformationPostfix(xyz, self.scaling_ * float(s))

    def scaledTransformVisiblePointsOfKeyframe(self, point_not_transformed_flags, diff_pose, kf_world_view_transform, kf_full_proj_transform,
                                               kf_creation_iter, stable_num_iter_existence, num_transformed=0, scale=1.0):
        """:420-475: Gaussians younger than `stable_num_iter_existence` relative to the keyframe, not yet transformed and in front of it
        are moved by diff_pose (operate_points.cu:95-143); xyz and rotation (ACTIVATED, as in the reference) re-enter the optimizer with
        zero moments. Returns the updated num_transformed."""
        from .points import scaleAndTransformThenMarkVisiblePoints
        points = self.xyz_.clone()
        rots = self.getRotationActivation()
        unstable = (self._exist() - int(kf_creation_iter)).abs() < int(stable_num_iter_existence)
        num_transformed = scaleAndTransformThenMarkVisiblePoints(points, rots, point_not_transformed_flags, unstable, diff_pose,
                                                                 kf_world_view_transform, kf_full_proj_transform, num_transformed, scale)
        self.xyz_, self.rotation_ = points.contiguous(), rots.contiguous()
        for i in (0, 5):
            self.exp_avg_[i] = torch.zeros_like(self.tensors()[i])
            self.exp_avg_sq_[i] = torch.zeros_like(self.tensors()[i])
        return num_transformed

    def _cmodel(self, with_state=True):
        m = _Model()
        for i, t in enumerate(self.tensors()):
            assert t.is_contiguous() and t.dtype == torch.float32
            m.param[i] = t.data_ptr()
            if with_state:
                m.exp_avg[i] = self.exp_avg_[i].data_ptr()
                m.exp_avg_sq[i] = self.exp_avg_sq_[i].data_ptr()
        if with_state:
            m.max_radii2D = self.max_radii2D_.data_ptr()
            m.xyz_gradient_accum = self.xyz_gradient_accum_.data_ptr()
            m.denom = self.denom_.data_ptr()
        return m


def _ccamera(cam):
    c = _Camera()
    c.viewmatrix, c.projmatrix, c.campos = cam["viewmatrix"].data_ptr(), cam["projmatrix"].data_ptr(), cam["campos"].data_ptr()
    c.tan_fovx, c.tan_fovy, c.width, c.height = float(cam["tanfovx"]), float(cam["tanfovy"]), int(cam["W"]), int(cam["H"])
    return c


class GaussianTrainer:
    """One fused training iteration per call. `cam` = dict(viewmatrix, projmatrix, campos: device tensors; tanfovx, tanfovy, W, H)."""

    def __init__(self, model: GaussianModel, opt: GaussianOptimizationParams = None, background=None):
        self.L = _bind()
        self.model = model
        self.opt = opt or GaussianOptimizationParams()
        h = C.c_void_p()
        _lib.check(self.L.psb_trainer_create(C.byref(h)), "psb_trainer_create")
        self.h = h
        self.background = background if background is not None else torch.zeros(3, device=model.device)
        self.iteration = 0

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.psb_trainer_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _cstep(self, densify_stats=True):
        m = self.model
        s = _Step()
        for i in range(6):
            s.lr[i] = m.lr_[i]
        s.beta1, s.beta2, s.eps = 0.9, 0.999, 1e-15
        s.step = m.step_ + 1
        s.lambda_dssim = self.opt.lambda_dssim
        s.sh_degree = m.active_sh_degree_
        s.update_densify_stats = int(densify_stats)
        return s

    def render(self, cam, out=None, radii=None, check=True):
        """Forward only. check=True waits for the instance count and renders again if the binning arena had to grow (an
        overflowing render leaves `out` blank); pass check=False to only enqueue (call result() yourself before using `out`)."""
        m = self.model
        out = out if out is not None else torch.empty((3, cam["H"], cam["W"]), device=m.device)
        cm, cc = m._cmodel(False), _ccamera(cam)
        _lib.check(self.L.psb_trainer_render(self.h, m.num_points(), 16, C.byref(cm), C.byref(cc), self.background.data_ptr(),
                                             m.active_sh_degree_, out.data_ptr(), radii.data_ptr() if radii is not None else None,
                                             torch.cuda.current_stream().cuda_stream), "psb_trainer_render")
        self._last_render = (cam, out, radii)
        self._last = None
        if check:
            self.result()
        return out

    def trainForOneIteration(self, cam, gt_image, mask=None, out_color=None, radii=None, densify_stats=None):
        """Enqueues render -> loss -> backward -> Adam on the current stream (no host sync)."""
        m = self.model
        self.iteration += 1
        if densify_stats is None:
            densify_stats = self.iteration < self.opt.densify_until_iter
        cm, cc, cs = m._cmodel(), _ccamera(cam), self._cstep(densify_stats)
        _lib.check(self.L.psb_trainer_step(self.h, m.num_points(), 16, C.byref(cm), C.byref(cc), self.background.data_ptr(),
                                           gt_image.data_ptr(), mask.data_ptr() if mask is not None else None, C.byref(cs),
                                           out_color.data_ptr() if out_color is not None else None,
                                           radii.data_ptr() if radii is not None else None,
                                           torch.cuda.current_stream().cuda_stream), "psb_trainer_step")
        m.step_ += 1
        self._last = (cam, gt_image, mask, out_color, radii, densify_stats)

    def _overflow_info(self):
        f, n, cur = C.c_uint(), C.c_uint(), C.c_uint()
        self.L.psb_trainer_overflow_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        self.L.psb_trainer_overflow_info(self.h, C.byref(f), C.byref(n), C.byref(cur))
        return f.value, n.value, cur.value

    def result(self):
        """Blocks; -> (loss, l1, ssim, num_rendered). Transparently repeats the call if the binning arena had to grow for it.
        The overflow record is sticky on the device: if an EARLIER queued step overflowed (it was a no-op on the model while
        step_ advanced), this raises instead of silently continuing."""
        out, n = (C.c_float * 3)(), C.c_int()
        rc = self.L.psb_trainer_result(self.h, out, C.byref(n), torch.cuda.current_stream().cuda_stream)
        if rc == -4:
            first, count, cur = self._overflow_info()
            if not (count == 1 and first == cur):
                if getattr(self, "_last", None) is not None:
                    self.model.step_ -= count      # those updates never happened
                    self.iteration -= count
                raise _lib.PsbError(f"{count} queued call(s) (first: pass {first} of {cur}) overflowed the binning arena and were dropped; the arena has "
                                    "been grown. Collect result() after every step, or render() the view once before queueing steps on it.")
            if getattr(self, "_last", None) is None:  # a render overflowed the arena: render again
                cam, o, rad = self._last_render
                self.render(cam, o, rad, check=False)
                return self.result()
            self.model.step_ -= 1                     # PSB_ERR_RETRY: the step was a no-op
            self.iteration -= 1
            cam, gt, mask, oc, rad, ds = self._last
            self.trainForOneIteration(cam, gt, mask, oc, rad, ds)
            return self.result()
        _lib.check(rc, "psb_trainer_result")
        return out[0], out[1], out[2], n.value

    # ------------------------------------------------------------------------------------------------------------
    # Host-input front end: what the mapper thread does every iteration — `gt_image = original_image_.cuda()`
    # (gaussian_mapper.cpp:637), train, `loss.item()` (:705). The copy of iteration i runs on a side stream while the
    # GPU still works on iteration i-1, and the loss comes back through psb_trainer_result's early read-back event
    # (recorded right behind the loss kernel), so the host enqueues iteration i+1 under the backward half of i.
    # ------------------------------------------------------------------------------------------------------------
    def trainHost(self, host_cam, host_gt, mask=None):
        """host_cam: dict(viewmatrix, projmatrix, campos: pinned CPU tensors; tanfovx, tanfovy, W, H); host_gt: pinned
        [3,H,W] CPU tensor. Enqueues this iteration and returns its loss (the GPU may still be in its backward half;
        flushHost() waits for it)."""
        dev = self.model.device
        if not hasattr(self, "_hs"):
            self._copy = torch.cuda.Stream()
            mk = lambda t: torch.empty(t.shape, dtype=t.dtype, device=dev)
            self._hs = [dict(gt=mk(host_gt), viewmatrix=mk(host_cam["viewmatrix"]), projmatrix=mk(host_cam["projmatrix"]),
                             campos=mk(host_cam["campos"]), ev=torch.cuda.Event(), done=torch.cuda.Event()) for _ in range(2)]
            self._hslot = 0
        slot = self._hs[self._hslot]
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self._copy):
            self._copy.wait_event(slot["done"])           # the iteration that last used this slot has finished reading it
            slot["gt"].copy_(host_gt, non_blocking=True)
            for k in ("viewmatrix", "projmatrix", "campos"):
                slot[k].copy_(host_cam[k], non_blocking=True)
            slot["ev"].record(self._copy)
        cam = dict(host_cam, viewmatrix=slot["viewmatrix"], projmatrix=slot["projmatrix"], campos=slot["campos"])
        main.wait_event(slot["ev"])
        self.trainForOneIteration(cam, slot["gt"], mask)
        loss = self.result()[0]                           # waits for the loss scalars only (repeats the step on RETRY)
        slot["done"].record(main)
        self._hslot ^= 1
        return loss

    def flushHost(self):
        """Waits until every enqueued iteration has finished on the device."""
        torch.cuda.current_stream().synchronize()

    STAGES = ("preprocess", "depth_sort_scan", "binning", "render_fwd", "loss", "render_bwd", "gaussian_backward", "frest_adam")
    DP_STAGES = ("preprocess", "depth_sort_scan", "binning", "render_fwd", "loss", "render_bwd", "push_backward", "wait_grads", "shard_adam", "wait_params",
                 "shard_adam_frest_part")

    def set_profiling(self, enable=True):
        self.L.psb_trainer_set_profiling.argtypes = [C.c_void_p, C.c_int]
        _lib.check(self.L.psb_trainer_set_profiling(self.h, int(enable)), "psb_trainer_set_profiling")

    def stage_times(self):
        """ms per stage of the last profiled step (CUDA events on the step's stream)."""
        ms = (C.c_float * 11)()
        self.L.psb_trainer_stage_times.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int]
        n = _lib.check(self.L.psb_trainer_stage_times(self.h, ms, 11), "psb_trainer_stage_times")
        names = self.DP_STAGES if n >= 10 else self.STAGES
        return dict(zip(names[:n], [float(x) for x in ms][:n]))

    def trainingOnce(self, cam, gt_image, mask=None):
        """Reference-style blocking iteration: returns the loss like loss.item() (gaussian_mapper.cpp:705)."""
        self.trainForOneIteration(cam, gt_image, mask)
        return self.result()[0]


class _DevArray:
    """Zero-copy view of library-owned device memory for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2, "strides": None}


class DataParallelTrainer(GaussianTrainer):
    """Keyframe-sharded data parallelism: replicated Gaussians, rank r renders its own view, K views per
    optimizer step (mean gradient: grad_scale = 1/world keeps the learning-rate scale of one view per step).

    mode="p2p" (default for world > 1): the fused NVLink step of psb_dp_step — the per-Gaussian backward pushes 80-byte
        gradient records straight into the inbox of the rank that owns the Gaussian (chunks of 128, owner = chunk % world),
        the owner applies Adam to its rows only and stores the updated rows into every rank's parameter tensors. No reduced
        gradient in memory, no collective call, optimizer traffic / world. The parameter tensors are re-homed in the
        context's IPC-mapped arena; Adam moments are valid on the owner rank only (gather_moments() before densification
        or a checkpoint).
    mode="nccl": round-1 path kept for A/B — slab-wise backward, one NCCL all-reduce per slab on a side stream, replicated Adam.
    world == 1 and mode=None: backward -> Adam split path on one GPU (tests)."""

    def __init__(self, model, opt=None, background=None, group=None, pipeline=None, nslabs=4, mode=None):
        super().__init__(model, opt, background)
        import os
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dropped_views = 0
        if mode is None:
            mode = os.environ.get("PSB_DP_MODE") or ("p2p" if self.world > 1 else "nccl")
        self.mode = mode
        self.dp = None
        if self.mode == "p2p":
            try:
                self._setup_p2p()
                ok = 1
            except _lib.PsbError as e:
                self._p2p_error = str(e)
                ok = 0
            if self.world > 1:   # every rank must take the same path
                t = torch.tensor([ok], device=model.device, dtype=torch.int32)
                dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
                ok = int(t.item())
            if not ok:
                import warnings
                warnings.warn("psb200: NVLink peer mapping unavailable (%s); falling back to the NCCL all-reduce path" % getattr(self, "_p2p_error", "a peer failed"))
                self._teardown_p2p()
                self.mode = "nccl"
        if self.mode == "nccl":
            self._setup_nccl(pipeline, nslabs)

    # ---- p2p -----------------------------------------------------------------------------------------------------------
    def _bind_dp(self):
        L, vp = self.L, C.c_void_p
        if getattr(L, "_dp_bound", False):
            return
        L.psb_dp_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int]
        L.psb_dp_ipc_handle.argtypes = [vp, vp]
        L.psb_dp_connect.argtypes = [vp, vp]
        L.psb_dp_params.argtypes = [vp, C.POINTER(vp)]
        L.psb_dp_step.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(_Model), C.POINTER(_Camera), vp, vp, vp, C.POINTER(_Step), vp, vp, vp]
        L.psb_dp_sync.argtypes = [vp, vp]
        L.psb_dp_status.argtypes = [vp, vp]
        L.psb_dp_destroy.argtypes = [vp]
        for n in ("psb_dp_create", "psb_dp_handle_bytes", "psb_dp_ipc_handle", "psb_dp_connect", "psb_dp_params", "psb_dp_step", "psb_dp_sync",
                  "psb_dp_status", "psb_dp_destroy"):
            getattr(L, n).restype = C.c_int
        L._dp_bound = True

    def _setup_p2p(self):
        from .parallel import SIZES
        self._bind_dp()
        m, L = self.model, self.L
        P = m.num_points()
        h = C.c_void_p()
        _lib.check(L.psb_dp_create(C.byref(h), self.rank, self.world, P), "psb_dp_create")
        self.dp = h
        nb = L.psb_dp_handle_bytes()
        if self.world > 1:
            mine = (C.c_char * nb)()
            _lib.check(L.psb_dp_ipc_handle(h, mine), "psb_dp_ipc_handle")
            blobs = [None] * self.world
            self.dist.all_gather_object(blobs, bytes(mine), group=self.group)
            allb = (C.c_char * (nb * self.world)).from_buffer_copy(b"".join(blobs))
            _lib.check(L.psb_dp_connect(h, allb), "psb_dp_connect")
        else:
            _lib.check(L.psb_dp_connect(h, None), "psb_dp_connect")
        ptrs = (C.c_void_p * 6)()
        _lib.check(L.psb_dp_params(h, ptrs), "psb_dp_params")
        shapes = [(P, 3), (P, 1, 3), (P, 15, 3), (P, 1), (P, 3), (P, 4)]
        homed = []
        for ptr, shp, src in zip(ptrs, shapes, m.tensors()):
            t = torch.as_tensor(_DevArray(ptr, shp), device=m.device) if P > 0 else torch.empty(shp, device=m.device)
            t.copy_(src)
            homed.append(t)
        self._homed = homed
        m._set(homed)
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.group)   # nobody pushes into an arena that is still being filled

    def _teardown_p2p(self):
        if self.dp is not None:
            m = self.model
            torch.cuda.synchronize()
            if self.world > 1:
                self.dist.barrier(group=self.group)
            if getattr(self, "_homed", None) is not None and m.tensors()[0] is self._homed[0]:
                m._set([t.clone() for t in m.tensors()])   # the arena is about to disappear
            self._homed = None
            self.L.psb_dp_destroy(self.dp)
            self.dp = None

    def close(self):
        self._teardown_p2p()

    def owner_mask(self):
        """bool [P]: rows whose Adam moments live on this rank (chunks of 128 Gaussians, owner = chunk % world)."""
        from .parallel import owner_of_rows
        return owner_of_rows(self.model.num_points(), self.world, self.model.device) == self.rank

    def sync(self):
        """Device-side wait until every rank's updated rows of the last step have landed in this rank's tensors."""
        if self.mode == "p2p":
            _lib.check(self.L.psb_dp_sync(self.dp, torch.cuda.current_stream().cuda_stream), "psb_dp_sync")

    def status(self):
        return self.L.psb_dp_status(self.dp, torch.cuda.current_stream().cuda_stream) if self.mode == "p2p" else 0

    def render(self, cam, out=None, radii=None, check=True):
        """Forward only, on the complete replica: the owner-side Adam of the last step runs on the context's own stream and the peers
        store their rows asynchronously, so the render is ordered behind both (device-side wait)."""
        self.sync()
        return super().render(cam, out, radii, check)

    def gather_moments(self):
        """Every rank receives the Adam moments of all rows (they are only maintained on the owner): call before
        densify/prune (which moves rows, hence owners) or before writing a checkpoint."""
        if self.mode != "p2p" or self.world == 1:
            return
        from .parallel import gather_owned_rows
        self.sync()
        m = self.model
        gather_owned_rows(m.exp_avg_ + m.exp_avg_sq_, self.rank, self.world, self.group)

    def rebuild(self):
        """After the model was resized (densify / prune / insertion): new arena, new peer mappings."""
        if self.mode == "p2p":
            self._teardown_p2p()
            self._setup_p2p()
        else:
            self._setup_nccl(self.pipeline, len(self.grads.slabs) if self.pipeline else 4)

    # ---- nccl (round-1 path) -------------------------------------------------------------------------------------------
    def _setup_nccl(self, pipeline, nslabs):
        from .parallel import GradBuffer, SlabGradBuffer
        model = self.model
        self.pipeline = (self.world > 1) if pipeline is None else bool(pipeline)
        P = model.num_points()
        if self.pipeline:
            self.grads = SlabGradBuffer(P, model.device, nslabs)
            self._comm = torch.cuda.Stream()
            n = len(self.grads.slabs)
            self._ev_k = [torch.cuda.Event() for _ in range(n)]
            self._ev_c = [torch.cuda.Event() for _ in range(n)]
            vp = C.c_void_p
            self.L.psb_trainer_backward_begin.argtypes = [vp, C.c_int, C.c_int, C.POINTER(_Model), C.POINTER(_Camera), vp, vp, vp, C.POINTER(_Step), vp, vp, vp]
            self.L.psb_trainer_backward_slab.argtypes = [vp, C.c_int, C.c_int, C.POINTER(_Model), C.POINTER(_Camera), C.POINTER(_Step), C.c_int, C.c_int,
                                                         C.POINTER(vp), vp]
            self.L.psb_adam_flat.argtypes = [C.c_size_t, vp, vp, vp, vp, C.c_float, C.POINTER(_Step), C.c_float, vp]
            for nme in ("psb_trainer_backward_begin", "psb_trainer_backward_slab", "psb_adam_flat"):
                getattr(self.L, nme).restype = C.c_int
        else:
            self.grads = GradBuffer(P, model.device)
            self.segs = self.grads.segments
        self.flat = self.grads.flat

    def trainForOneIteration(self, cam, gt_image, mask=None, out_color=None, radii=None, densify_stats=None):
        m = self.model
        self.iteration += 1
        if densify_stats is None:
            densify_stats = self.iteration < self.opt.densify_until_iter
        cm, cc, cs = m._cmodel(), _ccamera(cam), self._cstep(densify_stats)
        main = torch.cuda.current_stream()
        stream = main.cuda_stream
        P = m.num_points()
        args = (self.h, P, 16, C.byref(cm), C.byref(cc), self.background.data_ptr(), gt_image.data_ptr(),
                mask.data_ptr() if mask is not None else None, C.byref(cs), out_color.data_ptr() if out_color is not None else None,
                radii.data_ptr() if radii is not None else None)
        if self.mode == "p2p":
            _lib.check(self.L.psb_dp_step(self.h, self.dp, *args[1:], stream), "psb_dp_step")
        elif not self.pipeline:
            assert self.grads.P == P, "gradient buffer was built for a different model size: call rebuild() after densify/prune"
            ptrs = (C.c_void_p * 6)(*[s.data_ptr() for s in self.segs])
            _lib.check(self.L.psb_trainer_backward(*args, ptrs, stream), "psb_trainer_backward")
            scale = self.grads.all_reduce(self.group)   # ONE collective per step: 59 floats per Gaussian
            _lib.check(self.L.psb_adam_update(P, 16, C.byref(cm), ptrs, C.byref(cs), scale, stream), "psb_adam_update")
        else:
            from .parallel import SIZES
            assert self.grads.P == P, "gradient buffer was built for a different model size: call rebuild() after densify/prune"
            _lib.check(self.L.psb_trainer_backward_begin(*args, stream), "psb_trainer_backward_begin")
            scale = 1.0 / self.world
            for s, (first, count) in enumerate(self.grads.slabs):
                ptrs = (C.c_void_p * 6)(*self.grads.kernel_pointers(s))
                _lib.check(self.L.psb_trainer_backward_slab(self.h, P, 16, C.byref(cm), C.byref(cc), C.byref(cs), first, count, ptrs, stream),
                           "psb_trainer_backward_slab")
                if self.world > 1:
                    self._ev_k[s].record(main)
                    with torch.cuda.stream(self._comm):
                        self._comm.wait_event(self._ev_k[s])
                        self.dist.all_reduce(self.grads.region(s), op=self.dist.ReduceOp.SUM, group=self.group)
                        self._ev_c[s].record(self._comm)
            for s, (first, count) in enumerate(self.grads.slabs):
                if self.world > 1:
                    main.wait_event(self._ev_c[s])
                for i, (blk, k) in enumerate(zip(self.grads.blocks(s), SIZES)):
                    p, a, b = m.tensors()[i].view(-1), m.exp_avg_[i].view(-1), m.exp_avg_sq_[i].view(-1)
                    o = first * k
                    _lib.check(self.L.psb_adam_flat(count * k, p[o:].data_ptr(), a[o:].data_ptr(), b[o:].data_ptr(), blk.data_ptr(), m.lr_[i],
                                                    C.byref(cs), scale, stream), "psb_adam_flat")
        m.step_ += 1
        self._last = (cam, gt_image, mask, out_color, radii, densify_stats)

    def result(self):
        """-> (loss, l1, ssim, num_rendered) of this rank's view. A view whose binning arena overflowed contributed a zero
        gradient to the step (every rank still stepped, replicas identical); it is counted in `dropped_views`, the arena
        has been grown, and the step is NOT repeated (a one-rank retry would desynchronise the group)."""
        out, n = (C.c_float * 3)(), C.c_int()
        rc = self.L.psb_trainer_result(self.h, out, C.byref(n), torch.cuda.current_stream().cuda_stream)
        if rc == -4:
            self.dropped_views += 1
            return float("nan"), float("nan"), float("nan"), n.value
        _lib.check(rc, "psb_trainer_result")
        return out[0], out[1], out[2], n.value

    def sync_densify_stats(self):
        """Reduce the rank-local densification statistics (call right before densify/prune)."""
        from .parallel import reduce_densify_stats
        m = self.model
        reduce_densify_stats(m.max_radii2D_, m.xyz_gradient_accum_, m.denom_, self.group)


def fused_loss(image, gt, mask=None, lambda_dssim=0.2, want_grad=True):
    """-> (loss, l1, ssim, dL_dimage or None) through psb_loss."""
    L = _bind()
    _, H, W = image.shape
    grad = torch.empty_like(image) if want_grad else None
    out = (C.c_float * 3)()
    _lib.check(L.psb_loss(H, W, image.contiguous().data_ptr(), gt.contiguous().data_ptr(), mask.data_ptr() if mask is not None else None,
                          float(lambda_dssim), grad.data_ptr() if want_grad else None, out, torch.cuda.current_stream().cuda_stream), "psb_loss")
    return out[0], out[1], out[2], grad
