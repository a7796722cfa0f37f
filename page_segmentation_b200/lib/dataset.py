# API mirror: the class / field / function names and argument lists in this file follow ocr4all_pixel_classifier
# (https://github.com/ocr-d-modul-2-segmentierung/page-segmentation, (c) its authors, licensed Apache-2.0 OR
# GPL-3.0-or-later) so that it drops in for the reference; the arithmetic underneath is this repository's own
# (pcs_* calls into libpcseg_b200.so).
"""Mirror of ocr4all_pixel_classifier/lib/dataset.py for the prediction path:
SingleData (:17-29), Dataset (:32-41), prepare_images (:131-150) and
DatasetLoader (:153-208) and the directory listing `list_dataset` (:44-111) that
feeds it (dataset layout + per-page normalisation JSON).  The train/test split
helpers (:247-289) belong to training and are out of scope."""
from __future__ import annotations

import json
from dataclasses import dataclass
from typing import Any, Callable, List, Optional, Tuple

import numpy as np

import os

from .colors import ColorMap
from ..lazy import DeviceArray as _DeviceArray


@dataclass
class SingleData:
    image: np.ndarray = None
    binary: Optional[np.ndarray] = None
    orig_binary: Optional[np.ndarray] = None
    mask: np.ndarray = None
    image_path: Optional[str] = None
    binary_path: Optional[str] = None
    mask_path: Optional[str] = None
    line_height_px: Optional[int] = 1
    original_shape: Tuple[int, int] = None
    output_path: Optional[str] = None
    user_data: Any = None

    def __getattribute__(self, name):
        """Same fields as dataset.py:17-29.  The loader leaves `image` / `binary` / `orig_binary` on the device
        (lazy.DeviceArray) for the predictor and the output stage; a caller reading the attribute gets a real numpy
        array (one device-to-host copy at the first access), so code written against the reference - cv2 calls
        included - sees what it expects.  The device stages read the fields through lazy.peek."""
        v = object.__getattribute__(self, name)
        if type(v) is _DeviceArray:
            return v.to_host()
        return v


@dataclass
class Dataset:
    data: List[SingleData]
    color_map: ColorMap

    def __len__(self):
        return len(self.data)

    def __iter__(self):
        return self.data.__iter__()


def list_dataset(root_dir, line_height_px=None, binary_dir_="binary_images", images_dir_="images", masks_dir_="masks",
                 masks_postfix="", normalizations_dir="normalizations", verify_filenames=False):
    """dataset.py:44-111: one {binary_path, image_path, mask_path, line_height_px} record per page of a dataset
    directory; without a fixed `line_height_px` every page takes "char_height" from its normalisation JSON
    (<root>/<normalizations_dir>/*, what `compute-image-normalizations` / compute_char_height writes)."""
    def files_in(folder, keep):
        return [os.path.join(folder, name) for name in sorted(os.listdir(folder)) if keep(name)]

    dirs = {k: os.path.join(root_dir, v) for k, v in (("bin", binary_dir_), ("img", images_dir_), ("mask", masks_dir_))}
    for d in (root_dir, dirs["bin"], dirs["img"], dirs["mask"]):
        if not os.path.exists(d):
            raise Exception("Dataset dir does not exist at '%s'" % d)

    bins = files_in(dirs["bin"], lambda n: True)
    # images may live next to the masks: with a postfix, everything NOT carrying it is an image (:73)
    imgs = files_in(dirs["img"], (lambda n: not n.endswith(masks_postfix)) if masks_postfix else (lambda n: True))
    masks = files_in(dirs["mask"], lambda n: n.endswith(masks_postfix))

    base_names = None
    if verify_filenames:
        def by_stem(paths, postfix=None):
            if postfix:
                stripped = [p[:-len(postfix)] if p.endswith(postfix) else p for p in paths]
                return {os.path.basename(p).split('.')[0]: p + postfix for p in stripped}
            return {os.path.basename(p).split('.')[0]: p for p in paths}

        b, i, m = by_stem(bins), by_stem(imgs), by_stem(masks, masks_postfix)
        base_names = set(b.keys()).intersection(set(i.keys())).intersection(set(m.keys()))
        bins, imgs, masks = ([table.get(name) for name in base_names] for table in (b, i, m))

    if not line_height_px:
        norm_dir = os.path.join(root_dir, normalizations_dir)
        if not os.path.exists(norm_dir):
            raise Exception(f"Norm dir does not exist at '{norm_dir}'")
        norm_files = files_in(norm_dir, lambda n: True)
        if verify_filenames:
            norm_files = [f for f in norm_files if any(os.path.basename(f).startswith(s) for s in base_names)]
        heights = []
        for f in norm_files:
            with open(f, 'r') as fh:
                heights.append(json.load(fh)["char_height"])
        assert (len(heights) == len(masks))
    else:
        heights = [line_height_px] * len(masks)

    if not (len(bins) == len(imgs) == len(masks)):
        raise Exception("Mismatch in dataset files length: %d, %d, %d!" % (len(bins), len(imgs), len(masks)))
    return [{"binary_path": bp, "image_path": ip, "mask_path": mp, "line_height_px": lh}
            for bp, ip, mp, lh in zip(bins, imgs, masks, heights)]


def imread(path: str, as_gray: bool = True) -> np.ndarray:
    """Stand-in for ocr4all.files.imread (dataset.py:169): 8-bit grey page."""
    import cv2
    img = cv2.imread(path, cv2.IMREAD_GRAYSCALE if as_gray else cv2.IMREAD_COLOR)
    if img is None:
        raise FileNotFoundError(path)
    return img


def imread_bin(path: str, white_is_fg: bool = True) -> np.ndarray:
    """Stand-in for ocr4all.files.imread_bin (dataset.py:172): {0,255} page,
    paper = 255 (threshold at mid-grey; pylib's exact rule is unpinned)."""
    img = imread(path, True)
    return np.where(img > 127, 255, 0).astype(np.uint8)


def prepare_images(image: np.ndarray, binary: np.ndarray, target_line_height: int, line_height_px: int,
                   max_width: Optional[int] = None, keep_orig_bin=False):
    """dataset.py:131-150 (device kernel behind pcs_preprocess)."""
    from ..runtime import prepare_images_device
    return prepare_images_device(image, binary, target_line_height, line_height_px, max_width, keep_orig_bin)


class DatasetLoader:
    def __init__(self, target_line_height, color_map: ColorMap, prediction=False, max_width=None):
        self.target_line_height = target_line_height
        self.prediction = prediction
        self.color_map = color_map
        self.max_width = max_width

    def _page_job(self, entry: SingleData):
        """The page of one entry (dataset.py:160-172) as a staging job.  The reference derives the binary page from the
        'image' attribute as well (:172): an in-memory `image` IS the binary page; a file is read once and thresholded,
        and goes up as one plane when the scan is strictly two-level (imread_bin changes nothing then)."""
        from ..pipeline import PageJob
        if entry.image is not None:
            img = np.asarray(entry.image)           # (an entry that was loaded before: its host copy)
            grey, binary = (lambda a=img: a), None
        else:
            img = imread(entry.image_path, as_gray=True)
            bin_ = np.where(img > 127, 255, 0).astype(np.uint8)             # imread_bin of the same file
            grey = lambda a=img: a
            binary = None if np.array_equal(bin_, img) else (lambda a=bin_: a)
        if img.ndim != 2:
            raise ValueError("prepare_images expects 2-D image and binary of equal shape")
        if img.dtype != np.uint8:
            raise NotImplementedError("the device preprocess takes uint8 grey pages (as imread(as_gray=True) of 8-bit scans yields)")
        scale = self.target_line_height / entry.line_height_px
        return PageJob(grey, binary, img.shape[0], img.shape[1], scale, self.max_width), img

    def _finish_entry(self, entry: SingleData, job, img: np.ndarray, device: int) -> SingleData:
        from ..lazy import DeviceArray
        from ..pipeline import staged_fields
        image, binary = staged_fields(job, device)
        bin_src = job.binary() if job.binary is not None else img

        # (the closure must not hold `entry`: entry -> DeviceArray -> closure -> entry would keep the page's device tensors
        # alive until a pass of the cyclic collector -- 118 MB per 64 pages piling up between collections)
        tlh, lh, mw = self.target_line_height, entry.line_height_px, self.max_width

        def orig_binary():
            from ..runtime import prepare_images_tensors
            return prepare_images_tensors(img, bin_src, tlh, lh, mw, keep_orig_bin=True, device=device)[2]

        if not self.prediction:
            from .util import preserving_resize
            mask = entry.mask if entry.mask is not None else self.color_map.imread_labels(entry.mask_path)
            mask = preserving_resize(mask, image.shape)
            assert (mask.shape == image.shape)
            entry.mask = mask.astype(np.uint8)
        entry.binary = binary
        entry.orig_binary = DeviceArray(img.shape, np.uint8, orig_binary, device, pinned=False)
        entry.image = image
        entry.original_shape = img.shape
        return entry

    def load_images(self, dataset_file_entry: SingleData) -> SingleData:
        """dataset.py:160-191, including its quirk that the binary page is derived from `image` / `image_path`
        (attribute name 'image', :172).  prepare_images runs on the device; `image`, `binary` and `orig_binary` come back
        as DeviceArrays (lazy.py): ndarray stand-ins that stay on the GPU for the predictor and turn into host arrays
        the first time the caller looks at their contents."""
        from ..pipeline import stager
        from ..runtime import default_device
        device = default_device()
        job, img = self._page_job(dataset_file_entry)
        stager(device).submit([job])
        return self._finish_entry(dataset_file_entry, job, img, device)

    def load_data(self, all_dataset_files) -> Dataset:
        """dataset.py:193-198.  The reference fans pages out to a 12-process pool because decoding and rescaling are CPU
        work; here files are decoded by a thread pool, pages are staged to the device in chunks by a background stager
        (pipeline.PageStager: page-locked ring, one upload + one batched prepare_images per chunk) and this call returns
        without waiting for them -- the Predictor picks every chunk up when it is ready."""
        from ..pipeline import host_pool, stager
        from ..runtime import default_device
        device = default_device()
        entries = list(all_dataset_files)
        need_files = [e for e in entries if e.image is None]
        prepared = list(host_pool().map(self._page_job, entries)) if need_files else [self._page_job(e) for e in entries]
        stager(device).submit([job for job, _ in prepared])
        out = [self._finish_entry(e, job, img, device) for e, (job, img) in zip(entries, prepared)]
        return Dataset(out, self.color_map)

    def load_data_from_json(self, files, type) -> Dataset:
        """dataset.py:200-208."""
        all_files = []
        for f in files:
            if type == "all":
                all_files += [SingleData(**d) for t in ["train", "test", "eval"] for d in json.load(open(f, 'r'))[t]]
            else:
                all_files += [SingleData(**d) for d in json.load(open(f, 'r'))[type]]
        print(f"Loading {len(all_files)} data of type {type}")
        return self.load_data(all_files)
