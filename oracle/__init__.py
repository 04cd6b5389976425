"""CPU oracle for the A-TVSNet inference hot path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy (+ torch-CPU for the conv primitives) restatement of the
reference's TensorFlow-1.5 graph code for the path

    homography_warping.py  ->  model.py:build_cost_volume -> cost_volume_reasoning
    (cnn_wrapper/atvsnet.py:StackedUNet_prob) -> cost_volume_aggregation
    (network.py:attention_aggregation) -> output_conv -> prob2depth(_upsample)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package (``a-tvsnet_b200``)
never imports it and has no CPU fallback.

PARITY PIN STATUS
-----------------
The reference ships no tests, no golden tensors for any intermediate of this path
and its released checkpoint (model.zip) is absent; it needs Python 2.7 + TF 1.5, so it
cannot be run here as-is.  The oracle is pinned in two ways instead:

* ``tests/golden/*.npz`` were produced by executing the reference's OWN source files
  (``/root/reference/atvsnet/homography_warping.py``, ``model.py``,
  ``cnn_wrapper/{network,atvsnet}.py``) under Python 3 on top of a NumPy stand-in for
  the ``tensorflow`` primitives they call (``tests/golden/tf_shim.py``; generating
  script ``tests/golden/make_golden.py``).  The graph wiring, layouts, conventions and
  op order are therefore the reference's; the leaf arithmetic (matmul, gather_nd,
  conv3d SAME padding, batch-norm moments, softmax) is the shim's statement of TF
  semantics, not TF itself.
* geometry known-answer checks on the bundled ``example/`` cameras and ground-truth
  depth (see tests/test_oracle_geometry.py).

Because the leaf arithmetic is not TensorFlow's own, the pin is PARTIAL: wherever the
shim and the oracle share an assumption about a TF primitive, parity is unpinned.
"""
