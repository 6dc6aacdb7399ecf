"""pgmp_b200 -- B200-native (sm_100a) post-backbone grouping path.

Drop-in for the reference's two hot-path factories
(``src/graph_constructor/__init__.py:4-5`` and
``src/Models/MessagePassingNetwork/__init__.py:27-73``):

    from pgmp_b200.graph_constructor import get_graph_constructor
    from pgmp_b200.Models.MessagePassingNetwork import get_mpn_model
    from pgmp_b200.Utils import pred_to_person

All compute runs in hand-written CUDA kernels behind the C-ABI declared in
``include/pgmp.h`` (``csrc/`` -> ``libpgmp.so``); there is no CPU fallback.
"""

__version__ = "0.1.0"

from . import config  # noqa: F401
