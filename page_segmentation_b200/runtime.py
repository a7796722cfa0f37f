"""Host-side runtime: one `Context` per GPU, numpy <-> device staging helpers and
the page-batch engine used by the Predictor mirror, bench.py and the tests.

PyTorch is plumbing only (device tensors, pinned host memory, streams,
torch.distributed); all arithmetic is in libpcseg_b200.so.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from .synth import scaled_shape


_engine_tokens = __import__("itertools").count(1)


def default_device() -> int:
    return int(os.environ.get("LOCAL_RANK", os.environ.get("PCSEG_DEVICE", "0")))


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _native.PcsError("no CUDA device visible: the B200 hot path has no CPU fallback")
    return torch


def get_context(device: Optional[int] = None) -> _native.Context:
    torch = _torch()
    dev = default_device() if device is None else int(device)
    torch.cuda.set_device(dev)
    ctx = _native.context(dev)
    ctx.use_torch_stream()
    return ctx


def bind_to_gpu_numa_node(device: int) -> Optional[dict]:
    """Pins the calling process to the CPUs of the NUMA node its GPU hangs off (sysfs: the PCI device's `numa_node` and
    that node's `cpulist`), so that host buffers allocated afterwards - first touch - and the copy threads are local to
    the GPU's PCIe root.  With one process per GPU on a two-socket box this is what keeps eight host<->device streams
    from crossing the socket interconnect.  Returns what it did, or None when the topology is not exposed."""
    torch = _torch()
    try:
        p = torch.cuda.get_device_properties(device)
        addr = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{addr}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return {"pci": addr, "numa_node": node, "cpus": len(allowed)}
    except (OSError, ValueError, AttributeError):
        return None


def to_device_u8(arr: np.ndarray, device: int):
    torch = _torch()
    a = np.ascontiguousarray(arr)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    if a.dtype != np.uint8:
        if a.size and (a.min() < 0 or a.max() > 255):
            raise ValueError("values outside 0..255 cannot be staged as uint8")
        a = a.astype(np.uint8)
    return torch.from_numpy(a).to(f"cuda:{device}", non_blocking=False)


def to_host(t) -> np.ndarray:
    """Device tensor -> fresh numpy array the caller owns (the reference's functions return new arrays); pageable
    memory, through the process's page-locked bounce buffer (results_to_host)."""
    return results_to_host(t, site=None)[0]


# Page-sized results (probabilities, class map, masks) of the per-page API land in blocks of an explicit, bounded pool of
# page-locked memory (lazy.PinnedPool): a block comes back when the caller drops the array, so a caller that consumes a
# Prediction and moves on (the reference's front ends: predict -> write masks -> next page) copies at PCIe speed; a
# caller that keeps every result exhausts the pool's budget and gets ordinary pageable arrays, filled through one
# page-locked bounce buffer.  Datasets are kept by design (DatasetLoader.load_data returns all pages), so their
# host copies are pageable unless PCSEG_PINNED_LOAD=1.
_PIN_LOAD = os.environ.get("PCSEG_PINNED_LOAD", "0") == "1"      # page-by-page loaders that drop each page may opt in


def results_to_host(*tensors, site: Optional[str] = "predict"):
    """Device tensors (None allowed) -> fresh numpy arrays the caller owns, all copies issued before one
    synchronisation (lazy.tensors_to_host).  `site=None`: pageable destinations."""
    _torch()
    from .lazy import tensors_to_host
    return tensors_to_host(tensors, pinned=site is not None)


# ---------------------------------------------------------------------------
# single-page helpers behind the reference-named functions
# ---------------------------------------------------------------------------
def prepare_images_tensors(image: np.ndarray, binary: np.ndarray, target_line_height: int, line_height_px: int,
                           max_width: Optional[int] = None, keep_orig_bin: bool = False, device: Optional[int] = None):
    """dataset.py:131-150 on the GPU -> (d_image, d_binary, d_orig_binary | None) uint8 device tensors."""
    torch = _torch()
    ctx = get_context(device)
    dev = ctx.device
    image = np.asarray(image)
    binary = np.asarray(binary)
    if image.ndim != 2 or binary.shape != image.shape:
        raise ValueError("prepare_images expects 2-D image and binary of equal shape")
    if image.dtype != np.uint8:
        raise NotImplementedError("the device preprocess takes uint8 grey pages (as imread(as_gray=True) of 8-bit scans yields)")
    scale = target_line_height / line_height_px
    H, W = image.shape
    Hs, Ws = scaled_shape(H, W, scale)
    second = max_width is not None and max_width / Ws < 1.0            # dataset.py:139-141
    H1, W1 = Hs, Ws
    if second:
        Hs, Ws = scaled_shape(H1, W1, max_width / W1)
    same = binary is image or (binary.dtype == np.uint8 and np.shares_memory(binary, image))
    d_grey = to_device_u8(image, dev)
    d_bin = d_grey if same else to_device_u8(binary, dev)
    d_image = torch.empty((Hs, Ws), dtype=torch.uint8, device=d_grey.device)
    d_binary = torch.empty((Hs, Ws), dtype=torch.uint8, device=d_grey.device)
    d_orig = torch.empty((H, W), dtype=torch.uint8, device=d_grey.device) if keep_orig_bin else None
    from .pipeline import _preprocess_lock
    with _preprocess_lock:
        if second:
            ctx.preprocess_max_width(d_grey, d_bin, 1, H, W, H1, W1, Hs, Ws, d_image, d_binary, d_orig)
        else:
            ctx.preprocess(d_grey, d_bin, 1, H, W, Hs, Ws, d_image, d_binary, d_orig)
    return d_image, d_binary, d_orig


def prepare_images_device(image: np.ndarray, binary: np.ndarray, target_line_height: int, line_height_px: int,
                          max_width: Optional[int] = None, keep_orig_bin: bool = False, device: Optional[int] = None):
    """dataset.py:131-150 on the GPU; returns numpy uint8 arrays like the reference."""
    d_image, d_binary, d_orig = prepare_images_tensors(image, binary, target_line_height, line_height_px, max_width,
                                                       keep_orig_bin, device)
    img, bin_, orig = results_to_host(d_image, d_binary, d_orig, site="load" if _PIN_LOAD else None)
    if keep_orig_bin:
        return img, bin_, orig
    return img, bin_


def resize_nearest_plane(arr: np.ndarray, target_shape: Tuple[int, int], device: Optional[int] = None) -> np.ndarray:
    """util.py:21-29 preserving_resize for a uint8-representable plane."""
    torch = _torch()
    ctx = get_context(device)
    if arr.ndim != 2:
        raise ValueError("preserving_resize expects a 2-D plane")
    d_src = to_device_u8(arr, ctx.device)
    Ho, Wo = int(target_shape[0]), int(target_shape[1])
    d_dst = torch.empty((Ho, Wo), dtype=torch.uint8, device=d_src.device)
    ctx.resize_nearest(d_src, 1, arr.shape[0], arr.shape[1], d_dst, Ho, Wo)
    return to_host(d_dst)


def connected_components_with_stats(img: np.ndarray, device: Optional[int] = None):
    """cv2.connectedComponentsWithStats(img, connectivity=4) -> (n, labels i32, stats i32 (n,5))."""
    torch = _torch()
    ctx = get_context(device)
    H, W = img.shape
    d_img = to_device_u8((np.asarray(img) != 0), ctx.device)
    d_labels = torch.empty((H, W), dtype=torch.int32, device=d_img.device)
    d_ncomp = torch.zeros((1,), dtype=torch.int32, device=d_img.device)
    ctx.ccl(d_img, 1, H, W, d_labels, None, 0, d_ncomp)
    n = int(d_ncomp.cpu()[0])
    d_stats = torch.empty((n, 5), dtype=torch.int32, device=d_img.device)
    ctx.ccl(d_img, 1, H, W, d_labels, d_stats, n, d_ncomp)
    return n, d_labels.cpu().numpy(), d_stats.cpu().numpy()


# ---------------------------------------------------------------------------
# page-batch engine (device-resident or host-staged)
# ---------------------------------------------------------------------------
class PageBatchEngine:
    """Runs prepare_images -> network -> [cc_majority] -> masks for batches of
    equally sized pages on one GPU.  Buffers are allocated once per shape."""

    def __init__(self, arch: str, weights, n_classes: int, precision: str = "fp16", device: Optional[int] = None,
                 lut: Optional[np.ndarray] = None, engine: str = "umma"):
        self.torch = _torch()
        self.ctx = get_context(device)
        self._model = (arch, n_classes, list(weights), precision)
        self._engine = engine
        self._key = ("engine", next(_engine_tokens))
        self.n_classes = n_classes
        self._ensure_model()
        self.lut = None if lut is None else np.ascontiguousarray(lut, dtype=np.uint8)
        self._bufs: Dict[tuple, dict] = {}

    def _ensure_model(self):
        """The per-device context is shared and holds one model: reload ours if a Network or another engine has used
        the context since, and select our convolution engine (both are no-ops in the steady state)."""
        if self.ctx.loaded_key != self._key:
            arch, n_classes, weights, precision = self._model
            self.ctx.load_model(arch, n_classes, weights, precision, key=self._key)
        self.ctx.set_engine(self._engine)

    def buffers(self, n: int, H: int, W: int, Hs: int, Ws: int) -> dict:
        key = (n, H, W, Hs, Ws)
        b = self._bufs.get(key)
        if b is None:
            t, dev = self.torch, f"cuda:{self.ctx.device}"
            b = dict(
                image=t.empty((n, Hs, Ws), dtype=t.uint8, device=dev),
                binary=t.empty((n, Hs, Ws), dtype=t.uint8, device=dev),
                labels=t.empty((n, Hs, Ws), dtype=t.uint8, device=dev),
                color=t.empty((n, Hs, Ws, 3), dtype=t.uint8, device=dev),
                overlay=t.empty((n, Hs, Ws, 3), dtype=t.uint8, device=dev),
                inverted=t.empty((n, Hs, Ws, 3), dtype=t.uint8, device=dev),
            )
            self._bufs[key] = b
        return b

    def run_device(self, d_pages, scale: float, cc_majority: bool = False, masks: bool = True) -> dict:
        """d_pages: (n, H, W) uint8 CUDA tensor used as grey and binary page."""
        n, H, W = d_pages.shape
        Hs, Ws = scaled_shape(H, W, scale)
        b = self.buffers(n, H, W, Hs, Ws)
        ctx = self.ctx
        self._ensure_model()
        ctx.use_torch_stream()
        ctx.preprocess(d_pages, d_pages, n, H, W, Hs, Ws, b["image"], b["binary"], None)
        want = masks and self.lut is not None
        if cc_majority:
            ctx.forward(b["image"], b["binary"], n, Hs, Ws, b["labels"])
            ctx.cc_majority(b["labels"], b["binary"], n, Hs, Ws, self.n_classes)
            if want:
                ctx.masks(b["labels"], b["binary"], n, Hs, Ws, self.lut, b["color"], b["overlay"], b["inverted"])
        else:
            ctx.forward(b["image"], b["binary"], n, Hs, Ws, b["labels"], None, None,
                        self.lut if want else None, b["color"] if want else None,
                        b["overlay"] if want else None, b["inverted"] if want else None)
        return b

    def run_host(self, h_pages: np.ndarray, scale: float, out: dict, cc_majority: bool = False):
        """h_pages: (n, H, W) uint8 host array (ideally pinned); `out` holds host
        arrays 'labels' and optionally 'color','overlay','inverted' (pinned)."""
        n, H, W = h_pages.shape
        Hs, Ws = scaled_shape(H, W, scale)
        self._ensure_model()
        self.ctx.use_torch_stream()
        self.ctx.predict_pages_host(h_pages, h_pages, n, H, W, Hs, Ws, cc_majority, self.lut,
                                    None, None, out.get("labels"), out.get("color"), out.get("overlay"),
                                    out.get("inverted"))
        return out

    def run_host_segments(self, h_pages: np.ndarray, scale: float, out: dict, max_components: int, cc_majority: bool = True):
        """run_host followed by segment extraction (pcs_predict_pages_segments, BASELINE configs[3]): `out` holds host
        arrays 'labels' (n, Hs, Ws) uint8, 'stats' (n, n_classes, max_components, 5) int32, optionally 'ncomp'
        (n, n_classes) int32 and the masks 'color' / 'overlay' / 'inverted'."""
        n, H, W = h_pages.shape
        Hs, Ws = scaled_shape(H, W, scale)
        self._ensure_model()
        self.ctx.use_torch_stream()
        self.ctx.predict_pages_segments(h_pages, h_pages, n, H, W, Hs, Ws, cc_majority, self.lut, out["labels"], out["stats"],
                                        max_components, out.get("ncomp"), out.get("color"), out.get("overlay"), out.get("inverted"))
        return out

    def run_host_segments_compact(self, h_pages: np.ndarray, scale: float, out: dict, max_components: int, cc_majority: bool = True):
        """pcs_predict_pages_segments_compact: the segment call with compact results -- `out` holds 'labels', 'stats',
        optionally 'ncomp' and 'binary_bits' (see run_host_compact); no mask crosses PCIe."""
        n, H, W = h_pages.shape
        Hs, Ws = scaled_shape(H, W, scale)
        self._ensure_model()
        self.ctx.use_torch_stream()
        self.ctx.predict_pages_segments_compact(h_pages, h_pages, n, H, W, Hs, Ws, cc_majority, out["labels"], out.get("binary_bits"),
                                                out["stats"], max_components, out.get("ncomp"))
        return out

    def run_host_compact(self, h_pages: np.ndarray, scale: float, out: dict, cc_majority: bool = False):
        """pcs_predict_pages_compact: uint8 pages in; `out` holds 'labels' (n, Hs, Ws) uint8 and optionally 'binary_bits'
        (n, ceil(Hs * Ws / 32)) uint32 -- `data.binary` bit-packed.  The colour masks are produced on request
        (masks_from_compact) instead of crossing PCIe for every page."""
        n, H, W = h_pages.shape
        Hs, Ws = scaled_shape(H, W, scale)
        self._ensure_model()
        self.ctx.use_torch_stream()
        self.ctx.predict_pages_compact(h_pages, h_pages, n, H, W, Hs, Ws, cc_majority, out["labels"], out.get("binary_bits"))
        return out

    def submit_host_compact(self, h_pages: np.ndarray, scale: float, out: dict, cc_majority: bool = False, max_components: int = 0) -> int:
        """Streaming form of run_host_compact (and, with `max_components` and out['stats'], of run_host_segments_compact):
        queues the call and returns its ticket; `wait(ticket)` returns when `out` holds the results.  A submit that follows
        a submit of the same shapes is chained onto it -- its upload runs under the kernels of the call before -- so a
        caller that feeds batch after batch (Predictor.predict is a generator, predictor.py:27-30) pays the fill and the
        drain of the pipeline once, not once per batch.  Keep `h_pages` and `out` untouched until the wait; rotate over
        two or more sets of buffers to keep a batch in flight while the one before is consumed."""
        n, H, W = h_pages.shape
        Hs, Ws = scaled_shape(H, W, scale)
        self._ensure_model()
        self.ctx.use_torch_stream()
        if max_components:
            return self.ctx.predict_pages_segments_compact_submit(h_pages, h_pages, n, H, W, Hs, Ws, cc_majority, out["labels"],
                                                                  out.get("binary_bits"), out["stats"], max_components, out.get("ncomp"))
        return self.ctx.predict_pages_compact_submit(h_pages, h_pages, n, H, W, Hs, Ws, cc_majority, out["labels"], out.get("binary_bits"))

    def submit_host_packed(self, h_bits: np.ndarray, level0: int, level1: int, H: int, W: int, scale: float, out: dict,
                           cc_majority: bool = False) -> int:
        """Streaming form of run_host_packed (see submit_host_compact)."""
        n = h_bits.shape[0]
        Hs, Ws = scaled_shape(H, W, scale)
        if h_bits.dtype != np.uint32 or h_bits.shape[1] != (H * W + 31) // 32:
            raise ValueError("h_bits must be (n, ceil(H * W / 32)) uint32 (pack_pages)")
        self._ensure_model()
        self.ctx.use_torch_stream()
        return self.ctx.predict_pages_packed_submit(h_bits, level0, level1, n, H, W, Hs, Ws, cc_majority, out["labels"], out.get("binary_bits"))

    def wait(self, ticket: int):
        self.ctx.wait_pages(ticket)

    def run_host_packed(self, h_bits: np.ndarray, level0: int, level1: int, H: int, W: int, scale: float, out: dict,
                        cc_majority: bool = False):
        """pcs_predict_pages_packed: BIT-PACKED pages in (pack_pages), compact results out (see run_host_compact)."""
        n = h_bits.shape[0]
        Hs, Ws = scaled_shape(H, W, scale)
        if h_bits.dtype != np.uint32 or h_bits.shape[1] != (H * W + 31) // 32:
            raise ValueError("h_bits must be (n, ceil(H * W / 32)) uint32 (pack_pages)")
        self._ensure_model()
        self.ctx.use_torch_stream()
        self.ctx.predict_pages_packed(h_bits, level0, level1, n, H, W, Hs, Ws, cc_majority, out["labels"], out.get("binary_bits"))
        return out

    def masks_from_compact(self, labels: np.ndarray, binary_bits: np.ndarray) -> dict:
        """The colour masks of compact results, materialised on request ON THE DEVICE (pcs_unpack_bits + pcs_masks =
        generate_output_masks, output.py:44-60): {'color', 'overlay', 'inverted'} numpy (n, Hs, Ws, 3)."""
        t = self.torch
        dev = f"cuda:{self.ctx.device}"
        n, Hs, Ws = labels.shape
        self.ctx.use_torch_stream()
        d_labels = t.from_numpy(np.ascontiguousarray(labels)).to(dev)
        d_bits = t.from_numpy(np.ascontiguousarray(binary_bits).view(np.int32)).to(dev)
        d_bin = t.empty((n, Hs, Ws), dtype=t.uint8, device=dev)
        self.ctx.unpack_bits(d_bits, n, binary_bits.shape[1], Hs * Ws, d_bin)
        outs = t.empty((3, n, Hs, Ws, 3), dtype=t.uint8, device=dev)
        self.ctx.masks(d_labels, d_bin, n, Hs, Ws, self.lut, outs[0], outs[1], outs[2])
        color, overlay, inverted = results_to_host(outs[0], outs[1], outs[2], site="masks")
        return {"color": color, "overlay": overlay, "inverted": inverted}

    def run_host_files(self, h_pages: np.ndarray, scale: float, out: dict, cc_majority: bool = False):
        """Like run_host, but the three masks of every page come back as PNG files (pcs_predict_pages_files): `out` holds
        host arrays 'png' (n, 3, stride) uint8, 'png_sizes' (n, 3) uint64 and optionally 'labels' (ideally pinned).
        Returns `out`; file (p, k) is out['png'][p, k, :out['png_sizes'][p, k]], k = 0 color, 1 overlay, 2 inverted."""
        n, H, W = h_pages.shape
        Hs, Ws = scaled_shape(H, W, scale)
        self._ensure_model()
        self.ctx.use_torch_stream()
        self.ctx.predict_pages_files(h_pages, h_pages, n, H, W, Hs, Ws, cc_majority, self.lut, out.get("labels"), out["png"],
                                     out["png"].shape[2], out["png_sizes"])
        return out


def pack_pages(pages: np.ndarray):
    """Two-level uint8 pages (n, H, W) -> (bits (n, ceil(H * W / 32)) uint32, level0, level1): the layout
    pcs_preprocess_bits / pcs_predict_pages_packed take (flat, pixel i = bit i & 31 of word i >> 5; clear bit = level0).
    Host-side convenience for callers whose binarised pages are uint8 arrays; a loader that decodes 1-bit files
    (PBM, TIFF G4, 1-bit PNG) produces this form directly."""
    pages = np.ascontiguousarray(pages)
    levels = np.unique(pages)
    if pages.dtype != np.uint8 or pages.ndim != 3 or not 1 <= levels.size <= 2:
        raise ValueError("pack_pages expects uint8 pages (n, H, W) with at most two grey levels")
    level1 = int(levels[-1])
    level0 = int(levels[0]) if levels.size == 2 else (0 if level1 else 255)
    n = pages.shape[0]
    flat = (pages.reshape(n, -1) == level1)
    pad = (-flat.shape[1]) % 32
    if pad:
        flat = np.concatenate([flat, np.zeros((n, pad), dtype=bool)], axis=1)
    bits = np.packbits(flat, axis=1, bitorder="little").view("<u4")
    return np.ascontiguousarray(bits), level0, level1


def unpack_bits_host(bits: np.ndarray, shape) -> np.ndarray:
    """Inverse of the packed layout for one plane stack: (n, words) uint32 -> (n,) + shape uint8 {0,1} (host, tests)."""
    n = bits.shape[0]
    npix = int(np.prod(shape))
    flat = np.unpackbits(np.ascontiguousarray(bits).view(np.uint8).reshape(n, -1), axis=1, bitorder="little")[:, :npix]
    return flat.reshape((n,) + tuple(shape))


def shard_pages(n_pages: int, rank: int, world: int) -> List[int]:
    """Static round-robin page sharding over ranks (pages are independent;
    predictor.py:27-30 has batch 1 everywhere)."""
    return list(range(rank, n_pages, world))
