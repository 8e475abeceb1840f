"""CPU oracle for the AVLEN hot path (TEST INFRASTRUCTURE — NOT PRODUCT CODE).

Everything under ``oracle/`` is a CPU restatement of the reference's algorithm
for the hot path named in BASELINE.json.  It exists only so that
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` can check and time the reference
arithmetic.  The product package ``avlen_b200`` never imports it and fails
loudly if its CUDA library is missing.

Parity pinning:
  * model rows (C-K, N-Q): pinned against the reference's own PyTorch modules,
    imported unmodified from /root/reference under an import shim
    (``oracle/ref_shim.py``) in the authoring container; the resulting golden
    vectors live in ``tests/golden/`` next to ``make_golden.py``.
  * audio rows (A, B): ``scipy.signal.fftconvolve`` (the reference's real
    third-party code) plus a restatement of ``librosa.stft`` /
    ``skimage.measure.block_reduce`` (both absent from this image and unpinned
    in the reference's setup.py) -> **parity unpinned** at the librosa/skimage
    boundary; cross-checked three independent ways (numpy rfft frames,
    ``scipy.signal.stft``, ``torch.stft``) and against analytic known answers.
"""
