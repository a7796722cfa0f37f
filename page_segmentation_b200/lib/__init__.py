"""Drop-in mirror of `ocr4all_pixel_classifier.lib` for the inference hot path.

Module, class, function and field names follow the reference so that callers
only change the import root.  The arithmetic runs in
libpcseg_b200.so (hand-written sm_100a CUDA); nothing here falls back to CPU.
"""
