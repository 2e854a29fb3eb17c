"""CPU oracle for the SelfOcc hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain-PyTorch (CPU, fp32/fp64) restatement of the reference's
algorithm for the two hot paths (image->tri-plane lifting, SDF volume-render head).
It exists to *check* the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the product
package ``selfocc_b200`` never does (and fails loudly when its CUDA library is missing).

Provenance / pinning status (see DESIGN.md "Oracle"):

* PINNED by golden vectors generated from the importable pure-torch pieces of the
  reference (``tests/golden/make_golden.py``): grid<->metre mapping
  (model/encoder/bevformer/mappings.py), ``point_sampling``
  (model/encoder/bevformer/utils.py:116-206), ``get_cross_view_ref_points``
  (model/encoder/tpvformer/utils.py:5-71), ``RaySampler``
  (model/head/nerfacc_head/ray_sampler.py), SH bases (model/head/utils/sh_render.py),
  ``cal_depth_metric`` arithmetic (utils/metric_util.py:247-279).
* PARITY UNPINNED: the multi-scale deformable attention core (mmcv==2.0.1
  ``multi_scale_deform_attn``; not vendored, not installed) and the NeuS
  sampler/field/renderer (huang-yh/sdfstudio fork, unpinned, not vendored).  Their
  published semantics are restated from public knowledge and anchored on the
  reference's call sites; every such function says so in its docstring.
"""
