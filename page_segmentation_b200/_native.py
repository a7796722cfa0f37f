"""ctypes binding of libpcseg_b200.so (the C ABI in include/pcseg_b200.h).

There is deliberately no CPU or PyTorch fallback: if the shared library is
missing or the device is not an sm_100 GPU every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpcseg_b200.so")

ARCH_IDS = {"fcn_skip": 0, "fcn": 1, "unet": 2}
PRECISIONS = {"bf16": 0, "fp16": 1}
ENGINES = {"umma": 0, "direct": 1}

EXPORTS = [
    "pcs_abi_version", "pcs_ctx_create", "pcs_ctx_destroy", "pcs_last_error", "pcs_set_stream",
    "pcs_synchronize", "pcs_launch_count", "pcs_model_load", "pcs_set_engine", "pcs_preprocess",
    "pcs_preprocess_max_width", "pcs_preprocess_bits", "pcs_pack_bits", "pcs_unpack_bits", "pcs_predict_pages_compact", "pcs_predict_pages_packed",
    "pcs_forward", "pcs_masks", "pcs_resize_nearest", "pcs_ccl", "pcs_cc_majority",
    "pcs_bounding_boxes", "pcs_class_components", "pcs_char_height", "pcs_png_bytes", "pcs_png_encode", "pcs_output_pages", "pcs_output_flush", "pcs_segment_masks", "pcs_dilate3x3", "pcs_integral_image", "pcs_text_regions", "pcs_predict_pages_host", "pcs_predict_pages_files", "pcs_predict_pages_segments", "pcs_predict_pages_segments_compact", "pcs_predict_pages_compact_submit", "pcs_predict_pages_segments_compact_submit", "pcs_predict_pages_packed_submit", "pcs_wait_pages", "pcs_eval_counts",
    "pcs_train_input", "pcs_train_corr2d", "pcs_train_wgrad", "pcs_train_bias_grad", "pcs_train_relu_bwd", "pcs_train_maxpool_fwd",
    "pcs_train_maxpool_bwd", "pcs_train_deconv2_fwd", "pcs_train_deconv2_bwd_data", "pcs_train_deconv2_wgrad", "pcs_train_softmax_ce",
    "pcs_train_adam", "pcs_train_tc_create", "pcs_train_tc_step", "pcs_train_tc_destroy", "pcs_train_tc_wgrad", "pcs_set_saturation_check", "pcs_saturation_count", "pcs_debug_activation", "pcs_set_keep_activations", "pcs_set_timing", "pcs_set_pdl",
    "pcs_last_timings",
]


class PcsError(RuntimeError):
    pass


class LayerWeights(C.Structure):
    _fields_ = [("kernel", C.c_void_p), ("bias", C.c_void_p), ("shape", C.c_int32 * 4)]


_lib = None


def load() -> C.CDLL:
    """Load the in-tree shared library (built by page_segmentation_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PcsError(
            f"{LIB_PATH} is missing: run `python -m page_segmentation_b200.build` "
            "(there is no CPU fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, u8p = C.c_void_p, C.c_int, C.c_void_p
    lib.pcs_abi_version.restype = C.c_int
    lib.pcs_ctx_create.argtypes = [i32, C.POINTER(vp)]
    lib.pcs_ctx_destroy.argtypes = [vp]
    lib.pcs_ctx_destroy.restype = None
    lib.pcs_last_error.argtypes = [vp]
    lib.pcs_last_error.restype = C.c_char_p
    lib.pcs_set_stream.argtypes = [vp, vp]
    lib.pcs_synchronize.argtypes = [vp]
    lib.pcs_launch_count.argtypes = [vp]
    lib.pcs_launch_count.restype = C.c_int64
    lib.pcs_model_load.argtypes = [vp, i32, i32, i32, C.POINTER(LayerWeights), i32]
    lib.pcs_set_engine.argtypes = [vp, i32]
    lib.pcs_preprocess.argtypes = [vp, u8p, u8p, i32, i32, i32, i32, i32, u8p, u8p, u8p]
    lib.pcs_preprocess_max_width.argtypes = [vp, u8p, u8p, i32, i32, i32, i32, i32, i32, i32, u8p, u8p, u8p]
    lib.pcs_preprocess_bits.argtypes = [vp, vp, C.c_size_t, i32, i32, i32, i32, i32, i32, i32, u8p, u8p]
    lib.pcs_pack_bits.argtypes = [vp, u8p, i32, C.c_size_t, vp, C.c_size_t]
    lib.pcs_unpack_bits.argtypes = [vp, vp, i32, C.c_size_t, C.c_size_t, u8p]
    lib.pcs_predict_pages_compact.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.pcs_predict_pages_packed.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.pcs_forward.argtypes = [vp, u8p, u8p, i32, i32, i32, u8p, vp, vp, vp, u8p, u8p, u8p]
    lib.pcs_masks.argtypes = [vp, u8p, u8p, i32, i32, i32, vp, i32, u8p, u8p, u8p]
    lib.pcs_resize_nearest.argtypes = [vp, u8p, i32, i32, i32, u8p, i32, i32]
    lib.pcs_ccl.argtypes = [vp, u8p, i32, i32, i32, vp, vp, i32, vp]
    lib.pcs_cc_majority.argtypes = [vp, u8p, u8p, i32, i32, i32, i32]
    lib.pcs_bounding_boxes.argtypes = [vp, u8p, i32, i32, i32, i32, u8p]
    lib.pcs_char_height.argtypes = [vp, u8p, i32, i32, i32, i32, vp]
    lib.pcs_class_components.argtypes = [vp, u8p, i32, i32, i32, i32, vp, i32, vp]
    lib.pcs_png_bytes.argtypes = [i32, i32, i32, i32]
    lib.pcs_png_bytes.restype = C.c_size_t
    lib.pcs_png_encode.argtypes = [vp, u8p, i32, i32, i32, i32, i32, u8p, C.c_size_t, vp]
    lib.pcs_output_pages.argtypes = [vp, u8p, u8p, i32, i32, i32, vp, i32, C.POINTER(C.c_char_p)]
    lib.pcs_output_flush.argtypes = [vp]
    lib.pcs_segment_masks.argtypes = [vp, u8p, i32, i32, i32, i32, vp, i32, u8p]
    lib.pcs_dilate3x3.argtypes = [vp, u8p, i32, i32, i32, u8p]
    lib.pcs_integral_image.argtypes = [vp, u8p, i32, i32, i32, vp]
    lib.pcs_text_regions.argtypes = [vp, u8p, i32, i32, vp, i32, i32, i32, u8p, u8p]
    lib.pcs_predict_pages_host.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.pcs_predict_pages_files.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, C.c_size_t, vp]
    lib.pcs_predict_pages_segments.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp]
    lib.pcs_predict_pages_segments_compact.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, i32, vp]
    lib.pcs_predict_pages_compact_submit.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.pcs_predict_pages_segments_compact_submit.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, i32, vp, vp]
    lib.pcs_predict_pages_packed_submit.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.pcs_wait_pages.argtypes = [vp, C.c_uint64]
    lib.pcs_eval_counts.argtypes = [vp, vp, vp, vp, C.c_size_t, i32, vp]
    f32 = C.c_float
    lib.pcs_train_input.argtypes = [vp, vp, i32, i32, vp, i32, i32]
    lib.pcs_train_corr2d.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32]
    lib.pcs_train_wgrad.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32]
    lib.pcs_train_bias_grad.argtypes = [vp, vp, vp, i32, C.c_size_t]
    lib.pcs_train_relu_bwd.argtypes = [vp, vp, vp, C.c_size_t]
    lib.pcs_train_maxpool_fwd.argtypes = [vp, vp, vp, i32, i32, i32]
    lib.pcs_train_maxpool_bwd.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32]
    lib.pcs_train_deconv2_fwd.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32]
    lib.pcs_train_deconv2_bwd_data.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32]
    lib.pcs_train_deconv2_wgrad.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32]
    lib.pcs_train_softmax_ce.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp]
    lib.pcs_train_adam.argtypes = [vp, vp, vp, vp, vp, vp, i32, f32, f32, f32, f32, f32, f32]
    lib.pcs_train_tc_create.argtypes = [vp, i32, i32, i32, i32, vp, i32, C.POINTER(vp)]
    lib.pcs_train_tc_step.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp]
    lib.pcs_train_tc_destroy.argtypes = [vp, vp]
    lib.pcs_set_saturation_check.argtypes = [vp, i32]
    lib.pcs_saturation_count.argtypes = [vp, vp]
    lib.pcs_train_tc_wgrad.argtypes = [vp, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.pcs_debug_activation.argtypes = [vp, C.c_char_p, vp, C.c_size_t, C.POINTER(C.c_int32)]
    lib.pcs_set_keep_activations.argtypes = [vp, i32]
    lib.pcs_set_timing.argtypes = [vp, i32]
    lib.pcs_set_pdl.argtypes = [vp, i32]
    lib.pcs_last_timings.argtypes = [vp]
    lib.pcs_last_timings.restype = C.c_char_p
    if lib.pcs_abi_version() != 1:
        raise PcsError("libpcseg_b200.so ABI version mismatch")
    _lib = lib
    return lib


def _lut256(lut, n_classes: Optional[int] = None):
    """Colour LUT as a (256, 3) uint8 array (zero rows past the caller's table): the library reads n_classes rows of
    it (pcs_forward, pcs_predict_pages_*), so a table shorter than the model's class count must not reach it."""
    if lut is None:
        return None
    a = np.ascontiguousarray(lut, dtype=np.uint8).reshape(-1, 3)
    if a.shape[0] > 256:
        raise PcsError(f"colour LUT has {a.shape[0]} rows (at most 256)")
    if n_classes is not None and a.shape[0] < n_classes:
        raise PcsError(f"colour LUT has {a.shape[0]} rows but the model predicts {n_classes} classes")
    out = np.zeros((256, 3), dtype=np.uint8)
    out[:a.shape[0]] = a
    return out


def _ptr(t) -> Optional[int]:
    """Device/host pointer of a torch tensor / numpy array / None."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        if not t.flags["C_CONTIGUOUS"]:
            raise PcsError("numpy buffer must be C-contiguous")
        return t.ctypes.data
    if not t.is_contiguous():
        raise PcsError("tensor must be contiguous")
    return t.data_ptr()


class Context:
    """One pcs_ctx: one GPU, one stream, one loaded model."""

    def __init__(self, device: int = 0):
        self.lib = load()
        self.device = int(device)
        h = C.c_void_p()
        rc = self.lib.pcs_ctx_create(self.device, C.byref(h))
        if rc != 0:
            raise PcsError(f"pcs_ctx_create(device={device}) failed with status {rc} "
                           "(needs a CUDA sm_100 / B200 device; no fallback exists)")
        self.h = h
        self.model: Optional[Tuple[str, int, str]] = None
        self.loaded_key = None
        self._keepalive = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.pcs_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.pcs_last_error(self.h)
            raise PcsError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    # -- plumbing ----------------------------------------------------------
    def set_stream(self, cuda_stream: int):
        self._check(self.lib.pcs_set_stream(self.h, C.c_void_p(cuda_stream)), "pcs_set_stream")

    def use_torch_stream(self):
        import torch
        self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def synchronize(self):
        self._check(self.lib.pcs_synchronize(self.h), "pcs_synchronize")

    def launch_count(self) -> int:
        return int(self.lib.pcs_launch_count(self.h))

    def set_engine(self, engine: str):
        self._check(self.lib.pcs_set_engine(self.h, ENGINES[engine]), "pcs_set_engine")

    def set_keep_activations(self, enabled: bool):
        """Diagnostics: also store the activations the fused kernels never write (fcn_skip conv2)."""
        self._check(self.lib.pcs_set_keep_activations(self.h, 1 if enabled else 0), "pcs_set_keep_activations")

    def set_pdl(self, enabled: bool):
        self._check(self.lib.pcs_set_pdl(self.h, 1 if enabled else 0), "pcs_set_pdl")

    def set_timing(self, enabled: bool):
        self._check(self.lib.pcs_set_timing(self.h, 1 if enabled else 0), "pcs_set_timing")

    def timings(self):
        s = self.lib.pcs_last_timings(self.h).decode()
        out = []
        for item in s.split(";"):
            if item:
                k, v = item.rsplit(":", 1)
                out.append((k, float(v)))
        return out

    # -- model ---------------------------------------------------------------
    def load_model(self, arch: str, n_classes: int, weights: Sequence[Tuple[np.ndarray, np.ndarray]],
                   precision: str = "fp16", key=None):
        """Uploads a model.  The context holds ONE model; `key` names its owner (a Network / PageBatchEngine token) and
        is what `loaded_key` answers afterwards, so that every owner can tell whether the context still holds its
        weights.  A load without a key invalidates every owner's cache."""
        self.loaded_key = None
        arr = (LayerWeights * len(weights))()
        keep = []
        for i, (k, b) in enumerate(weights):
            k = np.ascontiguousarray(k, dtype=np.float32)
            b = np.ascontiguousarray(b, dtype=np.float32)
            if k.ndim != 4:
                raise PcsError(f"layer {i}: kernel must be 4-D, got shape {k.shape}")
            keep.append((k, b))
            arr[i].kernel = k.ctypes.data
            arr[i].bias = b.ctypes.data
            for d in range(4):
                arr[i].shape[d] = k.shape[d]
            if b.ndim != 1:
                raise PcsError(f"layer {i}: bias must be 1-D")
        self._check(self.lib.pcs_model_load(self.h, ARCH_IDS[arch], int(n_classes), PRECISIONS[precision],
                                            arr, len(weights)), "pcs_model_load")
        self.model = (arch, int(n_classes), precision)
        self._sat_unchecked = precision == "fp16"
        self.loaded_key = key

    # -- stages (device pointers) ----------------------------------------------
    def preprocess(self, d_grey, d_bin, n, H, W, Hs, Ws, d_image, d_binary, d_orig_binary=None):
        self._check(self.lib.pcs_preprocess(self.h, _ptr(d_grey), _ptr(d_bin), n, H, W, Hs, Ws, _ptr(d_image),
                                            _ptr(d_binary), _ptr(d_orig_binary)), "pcs_preprocess")

    def preprocess_max_width(self, d_grey, d_bin, n, H, W, H1, W1, H2, W2, d_image, d_binary, d_orig_binary=None):
        self._check(self.lib.pcs_preprocess_max_width(self.h, _ptr(d_grey), _ptr(d_bin), n, H, W, H1, W1, H2, W2,
                                                      _ptr(d_image), _ptr(d_binary), _ptr(d_orig_binary)),
                    "pcs_preprocess_max_width")

    def preprocess_bits(self, d_bits, words_per_page, n, H, W, level0, level1, Hs, Ws, d_image, d_binary=None):
        self._check(self.lib.pcs_preprocess_bits(self.h, _ptr(d_bits), words_per_page, n, H, W, int(level0), int(level1), Hs, Ws,
                                                 _ptr(d_image), _ptr(d_binary)), "pcs_preprocess_bits")

    def pack_bits(self, d_src, n, n_pixels, d_bits, words_per_page):
        self._check(self.lib.pcs_pack_bits(self.h, _ptr(d_src), n, n_pixels, _ptr(d_bits), words_per_page), "pcs_pack_bits")

    def unpack_bits(self, d_bits, n, words_per_page, n_pixels, d_dst):
        self._check(self.lib.pcs_unpack_bits(self.h, _ptr(d_bits), n, words_per_page, n_pixels, _ptr(d_dst)), "pcs_unpack_bits")

    def forward(self, d_image, d_binary, n, Hs, Ws, d_labels, d_logits=None, d_prob=None, lut=None,
                d_color=None, d_overlay=None, d_inverted=None):
        lut_arr = _lut256(lut, self.model[1] if self.model else None)
        self._check(self.lib.pcs_forward(self.h, _ptr(d_image), _ptr(d_binary), n, Hs, Ws, _ptr(d_labels),
                                         _ptr(d_logits), _ptr(d_prob), _ptr(lut_arr), _ptr(d_color),
                                         _ptr(d_overlay), _ptr(d_inverted)), "pcs_forward")
        self._after_forward()

    def masks(self, d_labels, d_binary, n, H, W, lut, d_color, d_overlay, d_inverted):
        lut_arr = np.ascontiguousarray(lut, dtype=np.uint8)
        self._check(self.lib.pcs_masks(self.h, _ptr(d_labels), _ptr(d_binary), n, H, W, _ptr(lut_arr),
                                       lut_arr.shape[0], _ptr(d_color), _ptr(d_overlay), _ptr(d_inverted)),
                    "pcs_masks")

    def resize_nearest(self, d_src, n, H, W, d_dst, Ho, Wo):
        self._check(self.lib.pcs_resize_nearest(self.h, _ptr(d_src), n, H, W, _ptr(d_dst), Ho, Wo),
                    "pcs_resize_nearest")

    def ccl(self, d_img, n, H, W, d_labels, d_stats=None, max_components=0, d_ncomp=None):
        self._check(self.lib.pcs_ccl(self.h, _ptr(d_img), n, H, W, _ptr(d_labels), _ptr(d_stats),
                                     int(max_components), _ptr(d_ncomp)), "pcs_ccl")

    def cc_majority(self, d_pred, d_binary, n, H, W, n_classes):
        self._check(self.lib.pcs_cc_majority(self.h, _ptr(d_pred), _ptr(d_binary), n, H, W, n_classes),
                    "pcs_cc_majority")

    def bounding_boxes(self, d_pred, n, H, W, n_classes, d_out):
        self._check(self.lib.pcs_bounding_boxes(self.h, _ptr(d_pred), n, H, W, n_classes, _ptr(d_out)),
                    "pcs_bounding_boxes")

    def class_components(self, d_pred, n, H, W, n_classes, d_stats, max_components, d_ncomp=None):
        self._check(self.lib.pcs_class_components(self.h, _ptr(d_pred), n, H, W, n_classes, _ptr(d_stats), int(max_components),
                                                  _ptr(d_ncomp)), "pcs_class_components")

    def char_height(self, d_img, n, H, W, inverse, d_height):
        self._check(self.lib.pcs_char_height(self.h, _ptr(d_img), n, H, W, 1 if inverse else 0, _ptr(d_height)),
                    "pcs_char_height")

    # -- image files ------------------------------------------------------------
    def png_bytes(self, H, W, channels, level=1) -> int:
        return int(self.lib.pcs_png_bytes(H, W, channels, level))

    def png_encode(self, d_img, n, H, W, channels, d_out, stride, d_sizes=None, level=1):
        self._check(self.lib.pcs_png_encode(self.h, _ptr(d_img), n, H, W, channels, level, _ptr(d_out), stride, _ptr(d_sizes)),
                    "pcs_png_encode")

    def output_pages(self, d_labels, d_binary, n, H, W, lut, paths):
        """pcs_output_pages: masks + three PNG files per page, written by the library's worker threads; `paths` page-major
        (color, overlay, inverted per page)."""
        lut_arr = np.ascontiguousarray(lut, dtype=np.uint8).reshape(-1, 3)
        if len(paths) != 3 * n:
            raise PcsError(f"output_pages: {3 * n} paths expected, got {len(paths)}")
        arr = (C.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
        self._check(self.lib.pcs_output_pages(self.h, _ptr(d_labels), _ptr(d_binary), n, H, W, _ptr(lut_arr), lut_arr.shape[0], arr),
                    "pcs_output_pages")

    def output_flush(self):
        self._check(self.lib.pcs_output_flush(self.h), "pcs_output_flush")

    # -- region extraction ---------------------------------------------------
    def segment_masks(self, d_rgb, H, W, Ho, Wo, colours, d_masks):
        cols = np.ascontiguousarray(colours, dtype=np.uint8).reshape(-1, 3)
        self._check(self.lib.pcs_segment_masks(self.h, _ptr(d_rgb), H, W, Ho, Wo, _ptr(cols), cols.shape[0], _ptr(d_masks)),
                    "pcs_segment_masks")

    def dilate3x3(self, d_src, H, W, C, d_dst):
        self._check(self.lib.pcs_dilate3x3(self.h, _ptr(d_src), H, W, C, _ptr(d_dst)), "pcs_dilate3x3")

    def integral_image(self, d_mask, n, H, W, d_sat):
        self._check(self.lib.pcs_integral_image(self.h, _ptr(d_mask), n, H, W, _ptr(d_sat)), "pcs_integral_image")

    def text_regions(self, d_rgb, H, W, colour, k_close, k_open, k_region, d_text_inv, d_region):
        col = np.ascontiguousarray(colour, dtype=np.uint8).reshape(3)
        self._check(self.lib.pcs_text_regions(self.h, _ptr(d_rgb), H, W, _ptr(col), int(k_close), int(k_open), int(k_region),
                                              _ptr(d_text_inv), _ptr(d_region)), "pcs_text_regions")

    # -- whole pipeline, host buffers ----------------------------------------
    def predict_pages_host(self, h_grey, h_bin, n, H, W, Hs, Ws, cc_majority=False, lut=None, h_image=None,
                           h_binary=None, h_labels=None, h_color=None, h_overlay=None, h_inverted=None):
        lut_arr = _lut256(lut, self.model[1] if self.model else None)
        self._check(self.lib.pcs_predict_pages_host(
            self.h, _ptr(h_grey), _ptr(h_bin), n, H, W, Hs, Ws, 1 if cc_majority else 0, _ptr(lut_arr),
            _ptr(h_image), _ptr(h_binary), _ptr(h_labels), _ptr(h_color), _ptr(h_overlay), _ptr(h_inverted)),
            "pcs_predict_pages_host")
        self._after_forward()

    def predict_pages_compact(self, h_grey, h_bin, n, H, W, Hs, Ws, cc_majority, h_labels, h_binary_bits=None):
        self._check(self.lib.pcs_predict_pages_compact(self.h, _ptr(h_grey), _ptr(h_bin), n, H, W, Hs, Ws, 1 if cc_majority else 0,
                                                       _ptr(h_labels), _ptr(h_binary_bits)), "pcs_predict_pages_compact")
        self._after_forward()

    def predict_pages_packed(self, h_bits, level0, level1, n, H, W, Hs, Ws, cc_majority, h_labels, h_binary_bits=None):
        self._check(self.lib.pcs_predict_pages_packed(self.h, _ptr(h_bits), int(level0), int(level1), n, H, W, Hs, Ws,
                                                      1 if cc_majority else 0, _ptr(h_labels), _ptr(h_binary_bits)),
                    "pcs_predict_pages_packed")
        self._after_forward()

    def predict_pages_files(self, h_grey, h_bin, n, H, W, Hs, Ws, cc_majority, lut, h_labels, h_png, png_stride, h_png_sizes):
        lut_arr = _lut256(lut, self.model[1] if self.model else None)
        self._check(self.lib.pcs_predict_pages_files(
            self.h, _ptr(h_grey), _ptr(h_bin), n, H, W, Hs, Ws, 1 if cc_majority else 0, _ptr(lut_arr),
            _ptr(h_labels), _ptr(h_png), png_stride, _ptr(h_png_sizes)), "pcs_predict_pages_files")
        self._after_forward()

    def predict_pages_segments(self, h_grey, h_bin, n, H, W, Hs, Ws, cc_majority, lut, h_labels, h_stats, max_components,
                               h_ncomp=None, h_color=None, h_overlay=None, h_inverted=None):
        lut_arr = _lut256(lut, self.model[1] if self.model else None)
        self._check(self.lib.pcs_predict_pages_segments(
            self.h, _ptr(h_grey), _ptr(h_bin), n, H, W, Hs, Ws, 1 if cc_majority else 0, _ptr(lut_arr), _ptr(h_labels),
            _ptr(h_color), _ptr(h_overlay), _ptr(h_inverted), _ptr(h_stats), int(max_components), _ptr(h_ncomp)),
            "pcs_predict_pages_segments")
        self._after_forward()

    def predict_pages_segments_compact(self, h_grey, h_bin, n, H, W, Hs, Ws, cc_majority, h_labels, h_binary_bits, h_stats,
                                       max_components, h_ncomp=None):
        self._check(self.lib.pcs_predict_pages_segments_compact(
            self.h, _ptr(h_grey), _ptr(h_bin), n, H, W, Hs, Ws, 1 if cc_majority else 0, _ptr(h_labels), _ptr(h_binary_bits),
            _ptr(h_stats), int(max_components), _ptr(h_ncomp)), "pcs_predict_pages_segments_compact")
        self._after_forward()

    def predict_pages_compact_submit(self, h_grey, h_bin, n, H, W, Hs, Ws, cc_majority, h_labels, h_binary_bits=None) -> int:
        """Streaming form of predict_pages_compact: returns the call's ticket once the work is queued (wait_pages)."""
        ticket = C.c_uint64(0)
        self._check(self.lib.pcs_predict_pages_compact_submit(self.h, _ptr(h_grey), _ptr(h_bin), n, H, W, Hs, Ws, 1 if cc_majority else 0,
                                                              _ptr(h_labels), _ptr(h_binary_bits), C.byref(ticket)),
                    "pcs_predict_pages_compact_submit")
        self._after_forward()
        return int(ticket.value)

    def predict_pages_segments_compact_submit(self, h_grey, h_bin, n, H, W, Hs, Ws, cc_majority, h_labels, h_binary_bits, h_stats,
                                              max_components, h_ncomp=None) -> int:
        ticket = C.c_uint64(0)
        self._check(self.lib.pcs_predict_pages_segments_compact_submit(
            self.h, _ptr(h_grey), _ptr(h_bin), n, H, W, Hs, Ws, 1 if cc_majority else 0, _ptr(h_labels), _ptr(h_binary_bits),
            _ptr(h_stats), int(max_components), _ptr(h_ncomp), C.byref(ticket)), "pcs_predict_pages_segments_compact_submit")
        self._after_forward()
        return int(ticket.value)

    def predict_pages_packed_submit(self, h_bits, level0, level1, n, H, W, Hs, Ws, cc_majority, h_labels, h_binary_bits=None) -> int:
        ticket = C.c_uint64(0)
        self._check(self.lib.pcs_predict_pages_packed_submit(self.h, _ptr(h_bits), int(level0), int(level1), n, H, W, Hs, Ws,
                                                             1 if cc_majority else 0, _ptr(h_labels), _ptr(h_binary_bits), C.byref(ticket)),
                    "pcs_predict_pages_packed_submit")
        self._after_forward()
        return int(ticket.value)

    def wait_pages(self, ticket: int):
        """Blocks until the results of that submit are in the caller's host buffers."""
        self._check(self.lib.pcs_wait_pages(self.h, int(ticket)), "pcs_wait_pages")

    def eval_counts(self, d_pred, d_mask, d_bin, n_pixels, n_classes, d_out):
        self._check(self.lib.pcs_eval_counts(self.h, _ptr(d_pred), _ptr(d_mask), _ptr(d_bin), n_pixels, n_classes, _ptr(d_out)), "pcs_eval_counts")

    # -- training primitives (lib/trainer.py walks the graph) ------------------------
    def train_call(self, name: str, *args):
        """pcs_train_<name>(ctx, *args); tensors / arrays are passed as pointers."""
        fn = getattr(self.lib, "pcs_train_" + name)
        conv = [(_ptr(a) if (a is None or hasattr(a, "data_ptr") or isinstance(a, np.ndarray)) else a) for a in args]
        self._check(fn(self.h, *conv), "pcs_train_" + name)

    # -- training step on the tensor cores (csrc/train_tc.cu) -----------------------
    def train_tc_create(self, arch: str, n_classes: int, h: int, w: int, offsets) -> int:
        offs = np.ascontiguousarray(offsets, dtype=np.int64)
        handle = C.c_void_p()
        self._check(self.lib.pcs_train_tc_create(self.h, ARCH_IDS[arch], int(n_classes), int(h), int(w), offs.ctypes.data, int(offs.size),
                                                 C.byref(handle)), "pcs_train_tc_create")
        return handle.value

    def train_tc_step(self, handle: int, phases: int, d_image, d_labels, d_params, d_grads, d_loss):
        self._check(self.lib.pcs_train_tc_step(self.h, handle, int(phases), _ptr(d_image), _ptr(d_labels), _ptr(d_params), _ptr(d_grads),
                                               _ptr(d_loss)), "pcs_train_tc_step")

    def train_tc_destroy(self, handle: int):
        if handle and self.h:
            self.lib.pcs_train_tc_destroy(self.h, handle)

    # -- diagnostics -----------------------------------------------------------
    def _after_forward(self):
        """First forward after a model load: the library has scanned the stored fp16 activations for saturated values."""
        if getattr(self, "_sat_unchecked", False):
            self._sat_unchecked = False
            n = self.saturation_count()
            if n:
                import warnings
                warnings.warn(f"{n} fp16 activation values of this model saturated at +-65504 on the first page(s); its class maps "
                              "may differ from the reference's fp32 arithmetic.  Use precision='bf16' (Network(precision=...), "
                              "PCSEG_PRECISION=bf16).", RuntimeWarning, stacklevel=3)

    def set_saturation_check(self, mode: int):
        self._check(self.lib.pcs_set_saturation_check(self.h, int(mode)), "pcs_set_saturation_check")

    def saturation_count(self) -> int:
        """fp16 activation stores that hit the +-65504 bound since the model was loaded (pcseg_b200.h)"""
        out = C.c_uint64(0)
        self._check(self.lib.pcs_saturation_count(self.h, C.byref(out)), "pcs_saturation_count")
        return int(out.value)

    def debug_activation(self, name: str) -> np.ndarray:
        shape = (C.c_int32 * 4)()
        c = self.lib.pcs_debug_activation(self.h, name.encode(), None, 0, shape)
        if c < 0:
            self._check(c, "pcs_debug_activation")
        out = np.empty(tuple(shape), dtype=np.float32)
        c = self.lib.pcs_debug_activation(self.h, name.encode(), out.ctypes.data, out.size, shape)
        if c < 0:
            self._check(c, "pcs_debug_activation")
        return out


_contexts = {}


def context(device: int = 0) -> Context:
    """Process-wide Context per device (created on first use)."""
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = Context(device)
        _contexts[device] = ctx
    return ctx
