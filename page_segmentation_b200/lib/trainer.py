# API mirror: the class / field / function names and argument lists in this file follow ocr4all_pixel_classifier
# (https://github.com/ocr-d-modul-2-segmentierung/page-segmentation, (c) its authors, licensed Apache-2.0 OR
# GPL-3.0-or-later) so that it drops in for the reference; the arithmetic underneath is this repository's own
# (pcs_* calls into libpcseg_b200.so).
"""Training step of the FCN graphs on the device: mirror of ocr4all_pixel_classifier/lib/trainer.py (`TrainSettings`
:59-106, `Trainer` :109-159) over `Network.train_dataset` (lib/network.py:167-242) for what BASELINE configs[4] names:
one page per step (the reference's batch, network.py:151-161), mean sparse cross entropy from logits (metrics.py:8-9),
Adam with per-variable clipnorm (network.py:91-103), data-parallel gradient averaging over the ranks of
torch.distributed.

Two engines behind `FcnTrainStep` (same parameters, same flat fp32 gradient buffer, same Adam):
  "tensor" (default)  csrc/train_tc.cu: mixed precision on the tensor cores -- bf16 activations and activation
                      gradients, fp32 accumulation, fp32 master weights -- one library call per step
                      (pcs_train_tc_step); with several ranks the gradient all-reduce of the decoder half starts while
                      the encoder half is still in backward;
  "fp32"              csrc/train.cu: fp32 on the CUDA cores; this module walks the graph of lib/model.py:45-92
                      (fcn_skip) or :206-234 (fcn) over the pcs_train_* primitives.  The numerics reference of the
                      tensor engine and the engine for more than four classes.
Concatenations never copy: a skip tensor is allocated inside the buffer of the concatenation it feeds, and its gradient
is the matching slice of that buffer's gradient.

Out of scope (Keras training UX): callbacks, early stopping, LR plateau, TensorBoard, augmentation, the other
optimizers and losses of lib/architecture.py:71-90 / lib/metrics.py.
"""
from __future__ import annotations

import os
from typing import List, NamedTuple, Optional, Sequence, Tuple

import numpy as np

from .architecture import Architecture
from .dataset import Dataset


class TrainSettings(NamedTuple):
    """Field names and defaults of trainer.py:59-106 (fields of the Keras training UX are accepted and ignored)."""
    n_epoch: int
    n_classes: int
    l_rate: float
    train_data: Dataset
    validation_data: Optional[Dataset]
    display: int
    output_dir: str
    threads: int
    data_augmentation: bool = False
    data_augmentation_settings: object = None
    early_stopping_max_performance_drops: int = 10
    early_stopping_restore_best_weights: bool = True
    early_stopping_min_delta: float = 0.0
    reduce_lr_on_plateau: bool = True
    reduce_lr_plateau_factor: float = 0.5
    reduce_lr_min_lr: float = 0.000001
    model_name: str = 'model'
    model_suffix: str = '.h5'
    save_best_model_only: bool = True
    save_weights_only: bool = False
    architecture: Architecture = Architecture.FCN_SKIP
    loss: object = None
    monitor: object = None
    optimizer: object = None
    optimizer_norm_clipping: bool = True
    optimizer_norm_clip_value: float = 1.0
    optimizer_clipping: bool = False
    optimizer_clip_value: float = 1.0
    evaluation_data: Optional[Dataset] = None
    load: Optional[str] = None
    continue_training: bool = False
    compute_baseline: bool = False
    foreground_masks: bool = False
    tensorboard: bool = False
    image_dimension: int = 1
    gpu_allow_growth: bool = False


# ---------------------------------------------------------------------------
# weight layouts: Keras <-> the correlation form the kernels use
# ---------------------------------------------------------------------------
def to_internal(kind: str, kernel: np.ndarray) -> np.ndarray:
    """conv / logits (kh,kw,Ci,Co) -> w[Co][Ci][ky][kx]; stride-1 transposed conv (kh,kw,Co,Ci) -> the flipped kernel in
    the same form (Conv2DTranspose 'same' == correlation with the flipped kernel); 2x2 stride-2 -> k2[tap][Co][Ci]."""
    k = np.asarray(kernel, dtype=np.float32)
    if kind in ("conv", "logits"):
        return np.ascontiguousarray(k.transpose(3, 2, 0, 1))
    if kind == "deconv":
        return np.ascontiguousarray(k[::-1, ::-1].transpose(2, 3, 0, 1))
    return np.ascontiguousarray(k.reshape(4, k.shape[2], k.shape[3]))


def from_internal(kind: str, w: np.ndarray, k: int) -> np.ndarray:
    w = np.asarray(w)
    if kind in ("conv", "logits"):
        return np.ascontiguousarray(w.transpose(2, 3, 1, 0))
    if kind == "deconv":
        return np.ascontiguousarray(w.transpose(2, 3, 0, 1)[::-1, ::-1])
    return np.ascontiguousarray(w.reshape(2, 2, w.shape[1], w.shape[2]))


def bwd_data_weights(kind: str, w):
    """Weights of the correlation that maps dy to dx for a correlation layer with weights w[Co][Ci][k][k]:
    w'[Ci][Co][ky][kx] = w[Co][Ci][k-1-ky][k-1-kx] (torch tensor in, torch tensor out; layout plumbing only)."""
    return w.flip(2, 3).transpose(0, 1).contiguous()


class FcnTrainStep:
    """Forward, loss, backward and Adam for one page of one of the FCN graphs; parameters live on the device as one
    flat fp32 buffer (variables in Keras order: kernel, bias per layer)."""

    def __init__(self, arch: str, weights: Sequence[Tuple[np.ndarray, np.ndarray]], n_classes: int, l_rate: float = 1e-3,
                 clipnorm: Optional[float] = 1.0, device: Optional[int] = None, engine: Optional[str] = None):
        from .. import runtime
        from ..synth import layer_table
        if arch not in ("fcn_skip", "fcn"):
            raise NotImplementedError("the device training step covers the fcn_skip and fcn graphs")
        engine = engine or os.environ.get("PCSEG_TRAIN_ENGINE") or ("tensor" if n_classes <= 4 else "fp32")
        if engine not in ("tensor", "fp32"):
            raise ValueError(f"unknown training engine {engine!r} (tensor | fp32)")
        if engine == "tensor" and n_classes > 4:
            raise ValueError("the tensor-core training step covers up to 4 classes; use engine='fp32'")
        self.engine = engine
        self._tc, self._tc_shape = None, None
        self._slots, self._copy_stream, self._turn, self._last_slot = None, None, 0, None
        self.torch = runtime._torch()
        self.ctx = runtime.get_context(device)
        self.arch, self.n_classes = arch, int(n_classes)
        self.table = layer_table(arch, n_classes)
        self.lr, self.clipnorm = float(l_rate), (float(clipnorm) if clipnorm else 0.0)
        self.b1, self.b2, self.eps, self.t = 0.9, 0.999, 1e-7, 0
        t, dev = self.torch, f"cuda:{self.ctx.device}"
        flat, offs, self.slots = [], [0], {}
        for (name, kind, k, ci, co, act), (kw, kb) in zip(self.table, weights):
            wi = to_internal(kind, kw)
            for arr in (wi, np.asarray(kb, dtype=np.float32)):
                flat.append(arr.reshape(-1))
                offs.append(offs[-1] + arr.size)
            self.slots[name] = (len(offs) - 3, wi.shape)
        self.params = t.from_numpy(np.concatenate(flat)).to(dev)
        self.grads = t.zeros_like(self.params)
        self.m = t.zeros_like(self.params)
        self.v = t.zeros_like(self.params)
        self.offsets = offs
        self.d_offsets = t.tensor(offs, dtype=t.int64, device=dev)
        self.d_loss = t.zeros((1,), dtype=t.float64, device=dev)
        self._shape = None

    # -- parameter views ----------------------------------------------------
    def _var(self, buf, name, which):
        i, wshape = self.slots[name]
        a, b = self.offsets[i + which], self.offsets[i + which + 1]
        v = buf[a:b]
        return v.view(wshape) if which == 0 else v

    def weights(self) -> List[Tuple[np.ndarray, np.ndarray]]:
        """Current parameters in Keras order / layout."""
        return [(from_internal(kind, self._var(self.params, name, 0).cpu().numpy(), k), self._var(self.params, name, 1).cpu().numpy())
                for (name, kind, k, ci, co, act) in self.table]

    def gradients(self) -> List[Tuple[np.ndarray, np.ndarray]]:
        """Gradients of the last step in Keras order / layout (before clipping)."""
        return [(from_internal(kind, self._var(self.grads, name, 0).cpu().numpy(), k), self._var(self.grads, name, 1).cpu().numpy())
                for (name, kind, k, ci, co, act) in self.table]

    # -- buffers ----------------------------------------------------------------
    def _alloc(self, h: int, w: int):
        from ..synth import padded_shape
        H, W = padded_shape(h, w)
        if self._shape == (h, w):
            return
        t, dev, skip = self.torch, f"cuda:{self.ctx.device}", self.arch == "fcn_skip"
        z = lambda c, hh, ww: t.zeros((c, hh, ww), dtype=t.float32, device=dev)      # noqa: E731
        H2, W2, H4, W4, H8, W8 = H // 2, W // 2, H // 4, W // 4, H // 8, W // 8
        A = {}
        A["x"] = z(1, H, W)
        A["conv1"] = z(20, H, W)
        A["d5cat"] = z(50 if skip else 20, H, W)             # [deconv5 | conv2]
        A["conv2"] = A["d5cat"][20:50] if skip else z(30, H, W)
        A["pool2"] = z(30, H2, W2)
        A["d4cat"] = z(70 if skip else 30, H2, W2)           # [deconv4 | conv3]
        A["conv3"] = A["d4cat"][30:70] if skip else z(40, H2, W2)
        A["conv4"] = z(40, H2, W2)
        A["pool4"] = z(40, H4, W4)
        A["d3cat"] = z(100 if skip else 40, H4, W4)          # [deconv3 | conv5]
        A["conv5"] = A["d3cat"][40:100] if skip else z(60, H4, W4)
        A["d2cat"] = z(120 if skip else 60, H4, W4)          # [deconv2 | conv6]
        A["conv6"] = A["d2cat"][60:120] if skip else z(60, H4, W4)
        A["pool6"] = z(60, H8, W8)
        A["conv7"] = z(80, H8, W8)
        A["deconv1"] = z(80, H8, W8)
        A["deconv2"], A["deconv3"], A["deconv4"], A["deconv5"] = A["d2cat"][0:60], A["d3cat"][0:40], A["d4cat"][0:30], A["d5cat"][0:20]
        A["logits"] = z(self.n_classes, H, W)
        self.act = A
        self.gact = {k: t.zeros_like(v) for k, v in A.items() if k in ("x", "conv1", "d5cat", "pool2", "d4cat", "conv4", "pool4", "d3cat",
                                                                      "d2cat", "pool6", "conv7", "deconv1", "logits")}
        G = self.gact
        if skip:
            G["conv2"], G["conv3"], G["conv5"], G["conv6"] = G["d5cat"][20:50], G["d4cat"][30:70], G["d3cat"][40:100], G["d2cat"][60:120]
        else:
            G["conv2"], G["conv3"], G["conv5"], G["conv6"] = (t.zeros_like(A[k]) for k in ("conv2", "conv3", "conv5", "conv6"))
        G["deconv2"], G["deconv3"], G["deconv4"], G["deconv5"] = G["d2cat"][0:60], G["d3cat"][0:40], G["d4cat"][0:30], G["d5cat"][0:20]
        self._shape, self._padded = (h, w), (H, W)

    # -- one step ---------------------------------------------------------------
    def __del__(self):
        try:
            if getattr(self, "_tc", None):
                self.ctx.train_tc_destroy(self._tc)
                self._tc = None
        except Exception:
            pass

    def _upload_page(self, image_u8, labels_u8):
        t, dev = self.torch, f"cuda:{self.ctx.device}"
        image_u8, labels_u8 = np.asarray(image_u8), np.asarray(labels_u8)
        if image_u8.ndim != 2 or labels_u8.shape != image_u8.shape:
            raise ValueError(f"image {image_u8.shape} and labels {labels_u8.shape} must be equally sized 2-D arrays")
        if labels_u8.size and int(labels_u8.max()) >= self.n_classes:
            raise ValueError(f"label {int(labels_u8.max())} outside 0..{self.n_classes - 1}")
        if self.engine != "tensor":
            return t.from_numpy(np.ascontiguousarray(image_u8, dtype=np.uint8)).to(dev), t.from_numpy(np.ascontiguousarray(labels_u8, dtype=np.uint8)).to(dev)
        # Two page slots of page-locked host memory + device buffers, filled on a copy stream: a pageable copy on the compute
        # stream would make the host wait for the previous step's kernels before it could enqueue this one (measured: 2.1 ms
        # per step against 1.8 ms of kernels).  A slot is reused two steps later, after the step that read it.
        n = image_u8.size
        if self._slots is None or self._slots[0]["pin"].shape[1] < n:
            self._copy_stream = t.cuda.Stream(self.ctx.device)
            self._slots = [{"pin": t.empty((2, n), dtype=t.uint8, pin_memory=True), "dev": t.empty((2, n), dtype=t.uint8, device=dev),
                            "copied": None, "consumed": None} for _ in range(2)]
            self._turn = 0
        slot = self._slots[self._turn % 2]
        self._turn += 1
        if slot["copied"] is not None:
            slot["copied"].synchronize()                          # the page-locked block is free again
        pin = slot["pin"].numpy()
        np.copyto(pin[0, :n].reshape(image_u8.shape), image_u8, casting="unsafe")
        np.copyto(pin[1, :n].reshape(image_u8.shape), labels_u8, casting="unsafe")
        main = t.cuda.current_stream(self.ctx.device)
        with t.cuda.stream(self._copy_stream):
            if slot["consumed"] is not None:
                self._copy_stream.wait_event(slot["consumed"])    # the step that read this slot's device buffers has run
            slot["dev"][:, :n].copy_(slot["pin"][:, :n], non_blocking=True)
            slot["copied"] = t.cuda.Event()
            slot["copied"].record(self._copy_stream)
        main.wait_event(slot["copied"])
        self._last_slot = slot
        return slot["dev"][0, :n].view(image_u8.shape), slot["dev"][1, :n].view(image_u8.shape)

    def _tc_enqueue(self, image_u8, labels_u8, phases: int = 3, page=None):
        """pcs_train_tc_step for one page (uploaded here unless `page` = (d_img, d_lab) is given); no synchronisation."""
        h, w = image_u8.shape
        self.ctx.use_torch_stream()
        if self._tc is None or self._tc_shape != (h, w):
            if self._tc:
                self.ctx.train_tc_destroy(self._tc)
            self._tc = self.ctx.train_tc_create(self.arch, self.n_classes, h, w, self.offsets)
            self._tc_shape = (h, w)
        d_img, d_lab = page if page is not None else self._upload_page(image_u8, labels_u8)
        self.ctx.train_tc_step(self._tc, phases, d_img, d_lab, self.params, self.grads, self.d_loss)
        if self._last_slot is not None and (phases & 1):
            ev = self.torch.cuda.Event()
            ev.record(self.torch.cuda.current_stream(self.ctx.device))
            self._last_slot["consumed"] = ev
        return d_img, d_lab

    def forward_backward(self, image_u8: np.ndarray, labels_u8: np.ndarray) -> float:
        """Fills self.grads with d loss / d parameters for one page and returns the loss."""
        if self.engine == "tensor":
            h, w = image_u8.shape
            self._tc_enqueue(image_u8, labels_u8)
            return float(self.d_loss.cpu()[0]) / (h * w)
        t, c = self.torch, self.ctx.train_call
        h, w = image_u8.shape
        self._alloc(h, w)
        H, W = self._padded
        A, G, skip = self.act, self.gact, self.arch == "fcn_skip"
        dev = f"cuda:{self.ctx.device}"
        d_img = t.from_numpy(np.ascontiguousarray(image_u8, dtype=np.uint8)).to(dev)
        d_lab = t.from_numpy(np.ascontiguousarray(labels_u8, dtype=np.uint8)).to(dev)
        P = lambda name, which: self._var(self.params, name, which)      # noqa: E731
        dP = lambda name, which: self._var(self.grads, name, which)      # noqa: E731
        info = {n: (kind, k, ci, co, act == "relu") for (n, kind, k, ci, co, act) in self.table}

        def corr(name, src, dst):
            kind, k, ci, co, relu = info[name]
            c("corr2d", src, P(name, 0), P(name, 1), dst, ci, co, src.shape[1], src.shape[2], k, int(relu), 0)

        def up(name, src, dst):
            kind, k, ci, co, relu = info[name]
            c("deconv2_fwd", src, P(name, 0), P(name, 1), dst, ci, co, src.shape[1], src.shape[2], int(relu))

        def pool(src, dst):
            c("maxpool_fwd", src, dst, src.shape[0], src.shape[1], src.shape[2])

        # ---- forward (model.py:45-92); slices of the concatenation buffers are contiguous channel ranges
        c("input", d_img, h, w, A["x"], H, W)
        corr("conv1", A["x"], A["conv1"]); corr("conv2", A["conv1"], A["conv2"]); pool(A["conv2"], A["pool2"])
        corr("conv3", A["pool2"], A["conv3"]); corr("conv4", A["conv3"], A["conv4"]); pool(A["conv4"], A["pool4"])
        corr("conv5", A["pool4"], A["conv5"]); corr("conv6", A["conv5"], A["conv6"]); pool(A["conv6"], A["pool6"])
        corr("conv7", A["pool6"], A["conv7"]); corr("deconv1", A["conv7"], A["deconv1"])
        up("deconv2", A["deconv1"], A["deconv2"])
        corr("deconv3", A["d2cat"] if skip else A["deconv2"], A["deconv3"])
        up("deconv4", A["d3cat"] if skip else A["deconv3"], A["deconv4"])
        up("deconv5", A["d4cat"] if skip else A["deconv4"], A["deconv5"])
        last = A["d5cat"] if skip else A["deconv5"]
        corr("logits", last, A["logits"])
        # ---- loss (metrics.py:8-9) over the crop; gradient of the logits
        c("softmax_ce", A["logits"], d_lab, self.n_classes, H, W, h, w, G["logits"], self.d_loss)

        # ---- backward
        def corr_bwd(name, src, g_dst, g_src, accumulate=0):
            """dst = act(corr(src)): masks g_dst by the ReLU, fills the layer's gradients, writes (or adds) g_src."""
            kind, k, ci, co, relu = info[name]
            dst = A[name]
            if relu:
                c("relu_bwd", g_dst, dst, g_dst.numel())
            c("wgrad", src, g_dst, dP(name, 0), ci, co, src.shape[1], src.shape[2], k)
            c("bias_grad", g_dst, dP(name, 1), co, src.shape[1] * src.shape[2])
            if g_src is not None:
                c("corr2d", g_dst, bwd_data_weights(kind, P(name, 0)), None, g_src, co, ci, src.shape[1], src.shape[2], k, 0, accumulate)

        def up_bwd(name, src, g_dst, g_src):
            kind, k, ci, co, relu = info[name]
            if relu:
                c("relu_bwd", g_dst, A[name], g_dst.numel())
            c("deconv2_wgrad", src, g_dst, dP(name, 0), ci, co, src.shape[1], src.shape[2])
            c("bias_grad", g_dst, dP(name, 1), co, 4 * src.shape[1] * src.shape[2])
            c("deconv2_bwd_data", g_dst, P(name, 0), g_src, ci, co, src.shape[1], src.shape[2])

        def pool_bwd(src_name, g_pool, accumulate):
            src = A[src_name]
            c("maxpool_bwd", src, g_pool, G[src_name], src.shape[0], src.shape[1], src.shape[2], accumulate)

        corr_bwd("logits", last, G["logits"], G["d5cat"] if skip else G["deconv5"])
        up_bwd("deconv5", A["d4cat"] if skip else A["deconv4"], G["deconv5"], G["d4cat"] if skip else G["deconv4"])
        up_bwd("deconv4", A["d3cat"] if skip else A["deconv3"], G["deconv4"], G["d3cat"] if skip else G["deconv3"])
        corr_bwd("deconv3", A["d2cat"] if skip else A["deconv2"], G["deconv3"], G["d2cat"] if skip else G["deconv2"])
        up_bwd("deconv2", A["deconv1"], G["deconv2"], G["deconv1"])
        corr_bwd("deconv1", A["conv7"], G["deconv1"], G["conv7"])
        corr_bwd("conv7", A["pool6"], G["conv7"], G["pool6"])
        pool_bwd("conv6", G["pool6"], 1 if skip else 0)                   # joins the skip gradient already in G[conv6]
        corr_bwd("conv6", A["conv5"], G["conv6"], G["conv5"], accumulate=1 if skip else 0)
        corr_bwd("conv5", A["pool4"], G["conv5"], G["pool4"])
        pool_bwd("conv4", G["pool4"], 0)
        corr_bwd("conv4", A["conv3"], G["conv4"], G["conv3"], accumulate=1 if skip else 0)
        corr_bwd("conv3", A["pool2"], G["conv3"], G["pool2"])
        pool_bwd("conv2", G["pool2"], 1 if skip else 0)
        corr_bwd("conv2", A["conv1"], G["conv2"], G["conv1"])
        corr_bwd("conv1", A["x"], G["conv1"], None)
        return float(self.d_loss.cpu()[0]) / (h * w)

    def allreduce_gradients(self) -> float:
        """Data parallel: SUM of the flat gradient buffer over the ranks (NCCL); returns the 1/world scale that
        apply_gradients folds into the update.  No-op without an initialised process group."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.grads, op=dist.ReduceOp.SUM)
            return 1.0 / dist.get_world_size()
        return 1.0

    def apply_gradients(self, grad_scale: float = 1.0):
        self.t += 1
        lr_t = self.lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        self.ctx.train_call("adam", self.params, self.grads, self.m, self.v, self.d_offsets, len(self.offsets) - 1,
                            float(lr_t), self.b1, self.b2, self.eps, self.clipnorm, float(grad_scale))

    def describe(self) -> dict:
        if self.engine == "tensor":
            return {"engine": "tensor", "dtype": "bf16 activations / activation gradients, fp32 accumulation, fp32 master weights and Adam",
                    "kernels": "conv_umma_kernel (forward, input gradients), wgrad_tc_kernel (tcgen05, MN-major operands)"}
        return {"engine": "fp32", "dtype": "f32 (CUDA cores)"}

    def _decoder_start(self) -> int:
        """offset of the first variable the backward pass finishes in its first phase (deconv1 .. logits)"""
        return self.offsets[self.slots["deconv1"][0]]

    def step(self, image_u8: np.ndarray, labels_u8: np.ndarray, lazy_loss: bool = False):
        """One optimisation step on one page; returns the loss as a float, or (tensor engine, lazy_loss=True) as a 0-d
        device tensor that costs no synchronisation until somebody reads it."""
        import torch.distributed as dist
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        if self.engine == "tensor" and world > 1:
            # data parallel: the decoder half of the gradients is all-reduced (NCCL, its own stream) while the encoder
            # half is still being computed
            h, w = image_u8.shape
            cut = self._decoder_start()
            page = self._tc_enqueue(image_u8, labels_u8, phases=1)
            first = dist.all_reduce(self.grads[cut:], op=dist.ReduceOp.SUM, async_op=True)
            self._tc_enqueue(image_u8, labels_u8, phases=2, page=page)
            second = dist.all_reduce(self.grads[:cut], op=dist.ReduceOp.SUM, async_op=True)
            first.wait()
            second.wait()
            self.apply_gradients(1.0 / world)
            loss = self.d_loss[0] / (h * w)
            return loss if lazy_loss else float(loss.cpu())
        if self.engine == "tensor" and lazy_loss:
            h, w = image_u8.shape
            self._tc_enqueue(image_u8, labels_u8)
            self.apply_gradients(1.0)
            return self.d_loss[0] / (h * w)
        loss = self.forward_backward(image_u8, labels_u8)
        self.apply_gradients(self.allreduce_gradients())
        return loss


class Trainer:
    """trainer.py:109-159 for the FCN graphs: n_epoch passes over train_data, one page per step, shuffled per epoch
    (network.py:133-135), the model written to <output_dir>/<model_name><model_suffix> at the end."""

    def __init__(self, settings: TrainSettings):
        from ..synth import make_weights
        self.settings = settings
        arch = settings.architecture.value if isinstance(settings.architecture, Architecture) else str(settings.architecture)
        if len(settings.train_data) == 0 and settings.n_epoch > 0:
            raise Exception("No training files specified. Maybe set n_iter=0")
        if settings.load:
            from .network import Network
            weights = Network("Predict", n_classes=settings.n_classes, model_constructor=Architecture(arch), model=settings.load).model.weights
        else:
            weights = make_weights(arch, settings.n_classes, seed=0)          # Glorot-uniform like the Keras layers' default
        self.arch = arch
        self.engine = FcnTrainStep(arch, weights, settings.n_classes, l_rate=settings.l_rate,
                                   clipnorm=settings.optimizer_norm_clip_value if settings.optimizer_norm_clipping else None)
        self.losses: List[float] = []

    def train(self, callback=None) -> None:
        s = self.settings
        data = list(s.train_data.data)
        rng = np.random.default_rng(0)
        for _epoch in range(s.n_epoch):
            rng.shuffle(data)
            for d in data:
                mask = np.asarray(d.mask, dtype=np.uint8).copy()
                if s.foreground_masks:
                    mask[np.asarray(d.binary) != 1] = 0                          # network.py:148-149
                self.losses.append(self.engine.step(np.asarray(d.image), mask, lazy_loss=True))
            self.losses = [float(v) for v in self.losses]            # one synchronisation per epoch, not per page
        if s.output_dir:
            self.save(os.path.join(s.output_dir, s.model_name + s.model_suffix))

    def save(self, path: str):
        from . import h5
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        h5.write_keras_h5(path, self.engine.weights(), model_name=self.arch, weights_only=self.settings.save_weights_only)
