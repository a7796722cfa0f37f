"""Arrays of the per-page API that stay on the device until somebody looks at them, and the page-locked block pool
their host copies come from.

The reference's API hands numpy arrays from call to call (`SingleData.image/binary`, `Prediction.labels`,
`generate_output_masks(data, pred, ...)`).  With every stage on the GPU a page would cross PCIe four times per call
chain; a `DeviceArray` is what those fields hold instead: shape / dtype / ndim answer without touching the data, our
own stages take the device tensor, and anything else (`np.asarray`, indexing, arithmetic, any ndarray attribute)
materialises ONE host copy.  From that moment the host copy is the array (a caller may have written into it), and a
later device consumer uploads it again.

Host copies land in blocks of an explicit bounded pool of page-locked memory (`PinnedPool`): a block is handed out as
the numpy array's memory and returns to the pool when the array and all its views are garbage collected (a
`weakref.finalize` on the base array; numpy collapses view chains onto it).  A caller that keeps every result exhausts
the pool (`PCSEG_PINNED_POOL_MB`, default 1024) and gets ordinary pageable arrays filled through one page-locked
bounce buffer; nothing is guessed from timings.
"""
from __future__ import annotations

import os
import threading
import weakref
from typing import Callable, Dict, List, Optional

import numpy as np

_POOL_BYTES = int(float(os.environ.get("PCSEG_PINNED_POOL_MB", "1024")) * (1 << 20))
_ENABLED = os.environ.get("PCSEG_PINNED_RESULTS", "1") != "0"


class PinnedPool:
    """Bounded pool of page-locked blocks in power-of-two size classes.

    `alloc(nbytes)` -> (uint8 numpy base array over a block, or None when the budget is spent).  The block is recycled
    when the base array dies; arrays handed to callers are views of it.  `allocator(nbytes)` must return an object with
    `.numpy()` (a page-locked torch uint8 tensor); tests inject a fake."""

    def __init__(self, budget_bytes: int = _POOL_BYTES, allocator: Optional[Callable[[int], object]] = None):
        self.budget = int(budget_bytes)
        self.allocator = allocator
        self.lock = threading.Lock()
        self.free: Dict[int, List[object]] = {}
        self.total = 0               # bytes of every block this pool has created (free or handed out)
        self.outstanding = 0         # bytes handed out
        self.stats = {"hits": 0, "new": 0, "refused": 0}

    @staticmethod
    def size_class(nbytes: int) -> int:
        n = max(int(nbytes), 1 << 16)
        return 1 << (n - 1).bit_length()

    def _release(self, cls: int, blk) -> None:
        with self.lock:
            self.free.setdefault(cls, []).append(blk)
            self.outstanding -= cls

    def alloc(self, nbytes: int) -> Optional[np.ndarray]:
        cls = self.size_class(nbytes)
        with self.lock:
            lst = self.free.get(cls)
            if lst:
                blk = lst.pop()
                self.stats["hits"] += 1
            elif self.total + cls <= self.budget:
                blk = None
                self.total += cls                      # reserved before the (slow) allocation below
                self.stats["new"] += 1
            else:
                self.stats["refused"] += 1
                return None
            self.outstanding += cls
        if blk is None:
            try:
                blk = self.allocator(cls)
            except Exception:
                with self.lock:
                    self.total -= cls
                    self.outstanding -= cls
                raise
        base = blk.numpy()
        weakref.finalize(base, self._release, cls, blk)
        return base

    def trim(self) -> None:
        """drop the free blocks (their memory goes back to the allocator)"""
        with self.lock:
            for cls, lst in self.free.items():
                self.total -= cls * len(lst)
            self.free.clear()


_pool: Optional[PinnedPool] = None
_bounce = {"buf": None, "lock": threading.Lock()}


def pinned_pool() -> PinnedPool:
    global _pool
    if _pool is None:
        import torch
        _pool = PinnedPool(_POOL_BYTES, lambda n: torch.empty((n,), dtype=torch.uint8, pin_memory=True))
    return _pool


def _bounce_buffer(torch, nbytes: int):
    buf = _bounce["buf"]
    if buf is None or buf.numel() < nbytes:
        _bounce["buf"] = buf = torch.empty((max(nbytes * 5 // 4, 32 << 20),), dtype=torch.uint8, pin_memory=True)
    return buf


def tensors_to_host(tensors, pinned: bool = True) -> list:
    """Device tensors (None allowed) -> numpy arrays the caller owns; every copy is issued before ONE synchronisation of
    the current stream.  Destinations are pool blocks while the pool has room, else fresh pageable arrays filled
    through the process's page-locked bounce buffer (the driver's own pageable path manages ~2 GB/s into untouched
    memory)."""
    import torch
    outs: list = []
    stream = None
    late = []
    for t in tensors:
        if t is None:
            outs.append(None)
            continue
        if t.numel() < (1 << 16):
            outs.append(t.cpu().numpy())
            continue
        t = t.contiguous()
        stream = torch.cuda.current_stream(t.device)
        nbytes = t.numel() * t.element_size()
        base = pinned_pool().alloc(nbytes) if (pinned and _ENABLED) else None
        if base is not None:
            dst = base[:nbytes].view(np.dtype(str(t.dtype).replace("torch.", ""))).reshape(tuple(t.shape))
            torch.from_numpy(dst).copy_(t, non_blocking=True)
            outs.append(dst)
        else:
            late.append((len(outs), t))
            outs.append(None)
    if late:
        with _bounce["lock"]:
            sizes = [(t.numel() * t.element_size() + 255) // 256 * 256 for _, t in late]
            buf = _bounce_buffer(torch, sum(sizes))
            views, off = [], 0
            for (i, t), size in zip(late, sizes):
                v = buf[off:off + t.numel() * t.element_size()].view(t.dtype).view(t.shape)
                v.copy_(t, non_blocking=True)
                views.append((i, v))
                off += size
            stream.synchronize()
            for i, v in views:
                outs[i] = np.array(v.numpy())          # a copy in fresh memory: the bounce buffer is reused
    elif stream is not None:
        stream.synchronize()
    return outs


class DeviceArray(np.lib.mixins.NDArrayOperatorsMixin):
    """A numpy-array stand-in whose data lives on the GPU until the first host access (module docstring).

    `source()` -> device tensor of `shape` (its dtype may be narrower than the reported `dtype`: class maps are uint8
    on the device and int64 to the caller, like np.argmax's result); it may block on a background stage."""

    __array_priority__ = 100.0

    def __init__(self, shape, dtype, source: Callable[[], object], device: int, pinned: bool = True):
        self._shape = tuple(int(s) for s in shape)
        self._dtype = np.dtype(dtype)
        self._source = source
        self._tensor = None
        self._host: Optional[np.ndarray] = None
        self._device = int(device)
        self._pinned = pinned
        self._lock = threading.Lock()

    # -- answered without touching the data ---------------------------------------------------------------------
    @property
    def shape(self):
        return self._shape

    @property
    def dtype(self):
        return self._dtype

    @property
    def ndim(self):
        return len(self._shape)

    @property
    def size(self):
        return int(np.prod(self._shape, dtype=np.int64))

    @property
    def nbytes(self):
        return self.size * self._dtype.itemsize

    def __len__(self):
        if not self._shape:
            raise TypeError("len() of unsized object")
        return self._shape[0]

    @property
    def on_device(self) -> bool:
        return self._host is None

    def __repr__(self):
        where = "device" if self._host is None else "host"
        return f"DeviceArray(shape={self._shape}, dtype={self._dtype}, {where})"

    # -- the two ways out ------------------------------------------------------------------------------------------
    def device_tensor(self):
        """The data as a device tensor in its device dtype (uploads the host copy if a caller has taken one)."""
        import torch
        with self._lock:
            if self._host is not None:
                h = np.ascontiguousarray(self._host)
                if h.dtype == np.bool_:
                    h = h.view(np.uint8)
                return torch.from_numpy(h).to(f"cuda:{self._device}")
            if self._tensor is None:
                self._tensor = self._source()
                self._source = None
            return self._tensor

    def to_host(self) -> np.ndarray:
        with self._lock:
            if self._host is None:
                if self._tensor is None:
                    self._tensor = self._source()
                    self._source = None
                t = self._tensor
                want = self._dtype
                import torch
                tdt = np.dtype(str(t.dtype).replace("torch.", ""))
                if tdt != want:
                    t = t.to(getattr(torch, want.name))          # widened on the device, not by a host pass
                host = tensors_to_host([t], self._pinned)[0]
                self._host = host.reshape(self._shape)
                self._tensor = None                              # the host copy is the array from now on
            return self._host

    # -- numpy protocols -------------------------------------------------------------------------------------------
    def __array__(self, dtype=None, copy=None):
        a = self.to_host()
        if dtype is not None and np.dtype(dtype) != a.dtype:
            return a.astype(dtype)
        return a.copy() if copy else a

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        inputs = tuple(x.to_host() if isinstance(x, DeviceArray) else x for x in inputs)
        if "out" in kwargs:
            kwargs["out"] = tuple(x.to_host() if isinstance(x, DeviceArray) else x for x in kwargs["out"])
        return getattr(ufunc, method)(*inputs, **kwargs)

    def __array_function__(self, func, types, args, kwargs):
        def conv(x):
            if isinstance(x, DeviceArray):
                return x.to_host()
            if isinstance(x, (list, tuple)):
                return type(x)(conv(y) for y in x)
            return x
        return func(*conv(args), **{k: conv(v) for k, v in kwargs.items()})

    def __getitem__(self, idx):
        return self.to_host()[idx]

    def __setitem__(self, idx, value):
        self.to_host()[idx] = value

    def __iter__(self):
        return iter(self.to_host())

    def __bool__(self):
        return bool(self.to_host())

    def __getattr__(self, name):
        # everything else an ndarray has (astype, copy, max, T, reshape, tobytes, flags, ...)
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.to_host(), name)


def device_tensor_of(x, device: int):
    """numpy array / DeviceArray / torch tensor -> device tensor (uint8 for bool)."""
    import torch
    if isinstance(x, DeviceArray):
        return x.device_tensor()
    if isinstance(x, torch.Tensor):
        return x
    a = np.ascontiguousarray(x)
    if a.dtype == np.bool_:
        a = a.view(np.uint8)
    return torch.from_numpy(a).to(f"cuda:{device}")


def peek(obj, name: str):
    """Field of a SingleData WITHOUT the materialisation its attribute access performs (lib/dataset.py)."""
    return object.__getattribute__(obj, name)


def is_lazy(x) -> bool:
    return isinstance(x, DeviceArray) and x.on_device
