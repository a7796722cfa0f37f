"""CPU oracle for the OCR4All pixel-classifier inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in `page_segmentation_b200/` may import this
package; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline
/ `--impl reference` legs use it, and there only as the checker or as the timed
CPU arm - never as the product path.

PARITY, per stage:
  * PINNED by vectors the reference's OWN code produced in the build container
    (tests/golden/make_reference_golden.py imports the reference modules that need only numpy + cv2:
    xycut, pc_segmentation, postprocess, output, image_ops, dataset's dataclasses; tests/golden/ref_*.npz;
    tests/test_reference_pins.py, also live wherever /root/reference is mounted): do_xy_cut, find_segments,
    get_text_contours, vote_connected_component_class, generate_output_masks (masking logic),
    compute_char_height, list_dataset, the value types and the post-processor registry;
  * UNPINNED: prepare_images (scikit-image 0.17.2 resize / rescale), the network graphs (tensorflow 2.5), the
    Keras HDF5 container (h5py) and ColorMap (ocr4all-pylib) - see below.

The training step (oracle/train.py: torch autograd on the restated graph + a numpy restatement of Keras' Adam with
clipnorm) inherits the network's status: unpinned, checked against finite differences and a hand-computed Adam step.

For the unpinned stages: the reference (`ocr4all_pixel_classifier` 0.6.5) ships no
tests, golden vectors or fixtures, and its arithmetic lives in third-party
packages that are not installed here and cannot be (tensorflow==2.5.0,
scikit-image==0.17.2, ocr4all-pylib==0.2.6, h5py==3.1.0; SURVEY.md section
8(c)).  This oracle therefore restates those packages' published algorithms at
the reference's own call sites (each function cites the reference file:line it
follows) and is pinned by hand-computed known-answer tests plus cv2 4.13
(installed) as a live oracle for connected components.
"""
