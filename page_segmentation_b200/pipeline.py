"""The reference's page-by-page flow (`DatasetLoader.load_data` -> `Predictor.predict` -> `output_data`,
dataset.py:193-198, predictor.py:27-42, output.py:20-41) as an asynchronous device pipeline.

* `PageStager` -- what the reference's 12-process pool (dataset.py:195) does, on one GPU: worker threads bring pages
  into a ring of page-locked chunk buffers (file decode / memcpy; both release the GIL), ONE stager thread with its
  own `pcs_ctx` and stream uploads each chunk and runs `prepare_images` for it (pcs_preprocess, batched).  `load_data`
  returns at once; the `SingleData` fields are `DeviceArray`s that wait for their chunk only when somebody needs them.
* `predict_stream` -- `Predictor.predict` with look-ahead: consecutive equally-sized pages go through the network as one
  batch, the registry post-processors run batched on the device, chunk k+1 is enqueued before chunk k's first
  `Prediction` is yielded, and nothing synchronises unless the caller reads an array.
* `output_data` for `.png` targets is ONE library call per page (pcs_output_pages, csrc/output.cu): masks and the three
  PNG files are produced on the device right away, worker threads of the library copy only the bytes of the files and
  write them.  `flush_outputs()` (also registered with atexit) waits until every file is on disk;
  `PCSEG_OUTPUT_ASYNC=0` makes every call wait.

All kernel launches on the shared per-device context stay on the caller's thread; the background threads own their
own context / streams, so the `pcs_ctx` contract (one thread at a time per context) holds.
"""
from __future__ import annotations

import atexit
import os
import queue
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from .lazy import DeviceArray, device_tensor_of, peek
from .synth import scaled_shape

CHUNK_PAGES = int(os.environ.get("PCSEG_CHUNK_PAGES", "8"))
STAGER_SLOTS = int(os.environ.get("PCSEG_STAGER_SLOTS", "3"))
HOST_THREADS = int(os.environ.get("PCSEG_HOST_THREADS", str(max(2, min(8, (os.cpu_count() or 4) // 2)))))
OUTPUT_ASYNC = os.environ.get("PCSEG_OUTPUT_ASYNC", "1") != "0"

TRACE: list = []                          # (stage, seconds) of the background threads when PCSEG_TRACE_API=1 (tools/profile_api.py)
_TRACE_ON = os.environ.get("PCSEG_TRACE_API", "0") == "1"
_preprocess_lock = threading.Lock()      # the anti-aliasing weights live in one __constant__ bank per process


def _torch():
    from .runtime import _torch as t
    return t()


_pool_lock = threading.Lock()
_pool: Optional[ThreadPoolExecutor] = None


def host_pool() -> ThreadPoolExecutor:
    global _pool
    with _pool_lock:
        if _pool is None:
            _pool = ThreadPoolExecutor(max_workers=HOST_THREADS, thread_name_prefix="pcseg-host")
        return _pool


# ---------------------------------------------------------------------------------------------------------------------
# loading
# ---------------------------------------------------------------------------------------------------------------------
class PageJob:
    """One page on its way to the device: where its pixels come from and the geometry prepare_images will produce."""

    __slots__ = ("grey", "binary", "H", "W", "Hs", "Ws", "H1", "W1", "second", "same", "chunk", "index")

    def __init__(self, grey: Callable[[], np.ndarray], binary: Optional[Callable[[], np.ndarray]], H: int, W: int,
                 scale: float, max_width: Optional[int]):
        self.grey, self.binary = grey, binary           # binary None: the same array as grey (dataset.py:169-172)
        self.same = binary is None
        self.H, self.W = H, W
        self.H1, self.W1 = scaled_shape(H, W, scale)
        self.second = max_width is not None and max_width / self.W1 < 1.0
        self.Hs, self.Ws = scaled_shape(self.H1, self.W1, max_width / self.W1) if self.second else (self.H1, self.W1)
        self.chunk = None
        self.index = -1

    def key(self):
        return (self.H, self.W, self.H1, self.W1, self.Hs, self.Ws, self.same)


class Chunk:
    """Up to CHUNK_PAGES consecutive pages of one geometry; `image` / `binary` are (n, Hs, Ws) uint8 device tensors once
    `ready` is set, valid on any stream after `wait()`."""

    def __init__(self, jobs: List[PageJob], device: int):
        self.jobs = jobs
        self.device = device
        self.ready = threading.Event()
        self.error: Optional[BaseException] = None
        self.image = self.binary = None
        self.event = None
        for i, j in enumerate(jobs):
            j.chunk, j.index = self, i

    def wait(self):
        """Block until the chunk is staged, then order the current stream after the stager's work."""
        self.ready.wait()
        if self.error is not None:
            raise self.error
        torch = _torch()
        stream = torch.cuda.current_stream(self.device)
        stream.wait_event(self.event)
        self.image.record_stream(stream)
        self.binary.record_stream(stream)
        return self


class PageStager:
    """One per device (module docstring).  Two threads form a pipeline over a ring of page-locked chunk buffers: the
    `bring` thread fills buffer k+1 (pool of memcpy / decode workers) while the `launch` thread uploads buffer k and
    enqueues prepare_images for it on its own stream and context."""

    def __init__(self, device: int):
        torch = _torch()
        self.device = device
        self.torch = torch
        self.q: "queue.Queue[Optional[Chunk]]" = queue.Queue()
        self.brought: "queue.Queue" = queue.Queue()
        self.ctx = None
        self.stream = None
        self.slots = [{"buf": None, "event": None, "free": threading.Semaphore(1)} for _ in range(max(2, STAGER_SLOTS))]
        self.thread = threading.Thread(target=self._bring_loop, name=f"pcseg-bring-{device}", daemon=True)
        self.launcher = threading.Thread(target=self._launch_loop, name=f"pcseg-stager-{device}", daemon=True)
        self.thread.start()
        self.launcher.start()

    def submit(self, jobs: Sequence[PageJob]) -> List[Chunk]:
        chunks: List[Chunk] = []
        run: List[PageJob] = []
        for j in jobs:
            if run and (len(run) == CHUNK_PAGES or run[0].key() != j.key()):
                chunks.append(Chunk(run, self.device))
                run = []
            run.append(j)
        if run:
            chunks.append(Chunk(run, self.device))
        for c in chunks:
            self.q.put(c)
        return chunks

    # -- bring thread: pages -> page-locked chunk buffer -------------------------------------------------------------
    def _bring_loop(self):
        torch = self.torch
        torch.cuda.set_device(self.device)
        turn = 0
        while True:
            chunk = self.q.get()
            if chunk is None:
                self.brought.put(None)
                return
            slot = self.slots[turn % len(self.slots)]
            turn += 1
            slot["free"].acquire()                           # released by the launch thread once the upload is enqueued
            try:
                t0 = time.perf_counter()
                if slot["event"] is not None:
                    slot["event"].synchronize()              # ... and that upload has left the buffer
                j0 = chunk.jobs[0]
                n, H, W = len(chunk.jobs), j0.H, j0.W
                planes = 1 if j0.same else 2
                nbytes = n * planes * H * W
                if slot["buf"] is None or slot["buf"].numel() < nbytes:
                    slot["buf"] = torch.empty((nbytes,), dtype=torch.uint8, pin_memory=True)
                host = slot["buf"][:nbytes].numpy().reshape(planes, n, H, W)
                t1 = time.perf_counter()

                def bring(i):
                    job = chunk.jobs[i]
                    g = job.grey()
                    if g.shape != (H, W) or g.dtype != np.uint8:
                        raise ValueError(f"page {i} of the chunk is {g.dtype} {g.shape}, expected uint8 {(H, W)}")
                    np.copyto(host[0, i], g)
                    if planes == 2:
                        np.copyto(host[1, i], job.binary())

                list(host_pool().map(bring, range(n)))
                if _TRACE_ON:
                    TRACE.extend([("stager.slot", t1 - t0), ("stager.bring", time.perf_counter() - t1)])
                self.brought.put((chunk, slot, host))
            except BaseException as e:                       # handed to whoever waits for the chunk
                slot["free"].release()
                chunk.error = e
                chunk.ready.set()

    # -- launch thread: upload + prepare_images ------------------------------------------------------------------
    def _launch_loop(self):
        torch = self.torch
        torch.cuda.set_device(self.device)
        self.ctx = _native.Context(self.device)
        self.stream = torch.cuda.Stream(self.device)
        self.ctx.set_stream(self.stream.cuda_stream)
        while True:
            item = self.brought.get()
            if item is None:
                return
            chunk, slot, host = item
            try:
                self._launch(chunk, slot, host)
            except BaseException as e:
                chunk.error = e
            finally:
                slot["free"].release()
            chunk.ready.set()

    def _launch(self, chunk: Chunk, slot: dict, host: np.ndarray):
        torch = self.torch
        t2 = time.perf_counter()
        j0 = chunk.jobs[0]
        planes, n, H, W = host.shape
        dev = f"cuda:{self.device}"
        with torch.cuda.stream(self.stream):
            d_pages = torch.empty((planes, n, H, W), dtype=torch.uint8, device=dev)
            ta = time.perf_counter()
            d_pages.copy_(torch.from_numpy(host), non_blocking=True)
            tb = time.perf_counter()
            slot["event"] = torch.cuda.Event()
            slot["event"].record(self.stream)
            tb2 = time.perf_counter()
            out = torch.empty((2, n, j0.Hs, j0.Ws), dtype=torch.uint8, device=dev)
            image, binary = out[0], out[1]
            d_grey, d_bin = d_pages[0], d_pages[planes - 1]
            tc = time.perf_counter()
            with _preprocess_lock:
                if j0.second:
                    self.ctx.preprocess_max_width(d_grey, d_bin, n, H, W, j0.H1, j0.W1, j0.Hs, j0.Ws, image, binary, None)
                else:
                    self.ctx.preprocess(d_grey, d_bin, n, H, W, j0.Hs, j0.Ws, image, binary, None)
            ev = torch.cuda.Event()
            ev.record(self.stream)
            del d_pages                                        # same stream: the allocator may reuse it after the kernels
        chunk.image, chunk.binary, chunk.event = image, binary, ev
        chunk.jobs = ()                                        # job -> chunk -> job would keep the device tensors until a GC pass
        if _TRACE_ON:
            te = time.perf_counter()
            TRACE.extend([("stager.launch", te - t2), ("launch.alloc", ta - t2), ("launch.copy", tb - ta), ("launch.event", tb2 - tb), ("launch.alloc2", tc - tb2),
                          ("launch.preprocess", te - tc)])


_stagers: Dict[int, PageStager] = {}
_stagers_lock = threading.Lock()


def stager(device: int) -> PageStager:
    with _stagers_lock:
        s = _stagers.get(device)
        if s is None or not (s.thread.is_alive() and s.launcher.is_alive()):
            s = _stagers[device] = PageStager(device)
        return s


def staged_fields(job: PageJob, device: int) -> Tuple[DeviceArray, DeviceArray]:
    """(image, binary) DeviceArrays of a submitted job."""
    def field(name):
        def source():
            c = job.chunk.wait()
            return getattr(c, name)[job.index]
        return DeviceArray((job.Hs, job.Ws), np.uint8, source, device, pinned=False)
    return field("image"), field("binary")


# ---------------------------------------------------------------------------------------------------------------------
# prediction
# ---------------------------------------------------------------------------------------------------------------------
def _same_storage_batch(tensors):
    """(n, ...) view if the tensors are consecutive contiguous slices of one contiguous base tensor, else None."""
    t0 = tensors[0]
    base = getattr(t0, "_base", None)
    if base is None or not base.is_contiguous() or not t0.is_contiguous() or base.dtype != t0.dtype:
        return None
    step = t0.numel() * t0.element_size()
    p0 = t0.data_ptr()
    for i, t in enumerate(tensors):
        if getattr(t, "_base", None) is not base or t.data_ptr() != p0 + i * step or t.shape != t0.shape:
            return None
    first = (p0 - base.data_ptr()) // t0.element_size()
    return base.view(-1)[first:first + len(tensors) * t0.numel()].view((len(tensors),) + tuple(t0.shape))


def gather_batch(arrays, device: int):
    """list of equally shaped page arrays (DeviceArray / numpy) -> one (n, H, W) device tensor, without a copy when they
    are the pages of one staged chunk."""
    torch = _torch()
    ts = [device_tensor_of(a, device) for a in arrays]
    if len(ts) > 1 or getattr(ts[0], "_base", None) is not None:
        v = _same_storage_batch(ts)
        if v is not None:
            return v
    return torch.stack(ts) if len(ts) > 1 else ts[0][None].contiguous()


def known_postprocessors() -> dict:
    from .lib import postprocess as pp
    return {pp.vote_connected_component_class: "cc_majority", pp.add_bounding_boxes: "bounding_boxes"}


def predict_stream(predictor, pages) -> Iterable:
    """Predictor.predict (predictor.py:27-30).  Falls back to predict_single page by page for what the batched device
    path does not cover (high_res_output, foreign network objects); post-processors that are not the registry's own
    run on host arrays, as the reference calls them."""
    from .lib.predictor_data import Prediction
    net, settings = predictor.network, predictor.settings
    if settings.high_res_output or not hasattr(net, "_context"):
        for data in pages:
            yield predictor.predict_single(data)
        return
    known = known_postprocessors()
    procs = list(settings.post_process or [])
    torch = _torch()

    def chunks():
        run = []
        for data in pages:
            img = peek(data, "image")
            shp = img.shape if isinstance(img, DeviceArray) else tuple(np.shape(img))
            if run and (len(run) == CHUNK_PAGES or shp != run_shape[0]):
                yield run
                run = []
            if not run:
                run_shape = [shp]
            run.append(data)
        if run:
            yield run

    def launch(run):
        ctx = net._context()
        dev = ctx.device
        ctx.use_torch_stream()
        n = len(run)
        d_image = gather_batch([peek(d, "image") for d in run], dev)
        if d_image.dtype != torch.uint8 or d_image.dim() != 3:
            raise ValueError("data.image must be a 2-D uint8 array (DatasetLoader output)")
        _, h, w = d_image.shape
        d_labels = torch.empty((n, h, w), dtype=torch.uint8, device=d_image.device)
        ctx.forward(d_image, None, n, h, w, d_labels)
        host_from = None
        for k, proc in enumerate(procs):
            kind = known.get(proc)
            if kind is None:
                host_from = k
                break
            if kind == "cc_majority":
                d_bin = gather_batch([peek(d, "binary") for d in run], dev)
                ctx.cc_majority(d_labels, d_bin, n, h, w, net.n_classes)
            else:
                d_out = torch.empty_like(d_labels)
                ctx.bounding_boxes(d_labels, n, h, w, net.n_classes, d_out)
                d_labels = d_out
        out = []
        for i, data in enumerate(run):
            labels = DeviceArray((h, w), np.int64, (lambda t=d_labels[i]: t), dev)
            if host_from is not None:
                pred = labels.to_host()
                for proc in procs[host_from:]:
                    pred = proc(pred, data)
                labels = pred
            prob = DeviceArray((h, w, net.n_classes), np.float32, (lambda d=data: net._probabilities_device(d)), dev)
            out.append(Prediction(labels, prob, data))
        return out

    pending = None
    for run in chunks():
        nxt = launch(run)
        if pending is not None:
            yield from pending
        pending = nxt
    if pending is not None:
        yield from pending


# ---------------------------------------------------------------------------------------------------------------------
# output
# ---------------------------------------------------------------------------------------------------------------------
def flush_outputs():
    """Returns when every file handed to output_data so far is on disk (pcs_output_flush); raises what a background
    write failed with."""
    for ctx in list(_native._contexts.values()):
        if getattr(ctx, "h", None):
            ctx.output_flush()


def _flush_at_exit():
    try:
        flush_outputs()
    except Exception as e:                       # the interpreter is going down: report, do not raise
        import sys
        print(f"pcseg_b200: {e}", file=sys.stderr)


atexit.register(_flush_at_exit)
