"""CPU oracle for the ConAN message-passing backbone (TEST INFRASTRUCTURE ONLY).

This package is a plain-PyTorch / numpy CPU restatement of the algorithm that
the reference (duyhominhnguyen/conan-fgw) runs for its per-conformer backbone:

* ``oracle.radius``  - torch-cluster 1.6.1 ``radius_graph`` (CUDA truncation
  rule); reference call sites
  ``conan_fgw/src/model/graph_embeddings/schnet_no_sum.py:160,208,342`` and
  ``torch_geometric_visnet.py:331-337``.
* ``oracle.schnet``  - PyG 2.3.0 ``SchNet`` stack + ConAN's ``SchNetNoSum``
  heads (``schnet_no_sum.py:90-232``).
* ``oracle.visnet``  - restatement of the vendored ViSNet
  (``torch_geometric_visnet.py``) + ConAN wrapper (``visnet.py:82-158``).

Pinning status
--------------
* SchNet / radius graph: **parity unpinned**.  The arithmetic lives in
  torch-geometric 2.3.0 / torch-cluster 1.6.1, which are not vendored in the
  reference, are not installed here and cannot be installed (no network); the
  reference ships no tests or golden vectors for this path.  The restatement
  follows the published upstream algorithm as frozen in SURVEY.md Appendix A and
  is anchored on the reference's call sites; it is cross-checked by analytic
  known answers (tests/test_oracle_schnet.py) and an fp64 twin.
* ViSNet: **pinned** against the reference's own vendored file run in the build
  container through a 3-symbol PyG shim (``oracle/pyg_shim.py``); the outputs
  are committed as ``tests/golden/visnet_*.pt`` with the generating script
  ``tests/golden/make_golden.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product
(``conan-fgw_b200``) never does: it fails loudly when its CUDA library is
missing instead of falling back to anything in here.
"""

from .radius import radius_graph_ref, radius_interaction_graph_ref  # noqa: F401
