"""CPU oracle for the post-backbone grouping path -- TEST INFRASTRUCTURE ONLY.

This package is a plain-numpy restatement of the reference algorithm
(nibox/Pose-Estimation-with-Message-Passing-Networks) for the hot path named
in BASELINE.json: graph constructor -> message-passing network -> heads /
threshold -> person grouping.  Every function cites the reference file:line
it follows (paths relative to the reference root, ``CG.py`` =
``src/graph_constructor/ConstructGraph.py``).

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product
package never imports it and has no CPU fallback.

Parity status ("Oracle"):

* PINNED against the reference itself: candidate selection, edge index for
  ``fully``, node / edge features, every MPN variant in scope and the
  node-threshold / subgraph step.  ``tests/golden/make_golden.py`` imports the
  unmodified reference files from ``/root/reference`` (third-party packages
  that are not installable here are shimmed in ``tests/golden/ref_shims.py``)
  and writes the fixtures ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
  checks this oracle against them.
* PARITY UNPINNED for two third-party algorithms that are absent from the
  reference tree and from this image: the tie order of ``torch_cluster.knn``
  (``torch-cluster==1.5.4``, requirements.txt:63) and the native
  ``andres_graph_wrapper`` GAEC solver (no source, no version pin;
  correlation_clustering_utils.py:15).  Both follow the published algorithm
  with the tie rule written down in the function docstrings.
"""

from . import assemble, gc, mpn, grouping, refine  # noqa: F401
