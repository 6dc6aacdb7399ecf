"""The reference loader lives in ``oracle/ref_shims.py``; this alias keeps the fixture scripts' import."""
from oracle.ref_shims import *  # noqa: F401,F403
from oracle.ref_shims import load_reference, load_reference_grouping, ref_src  # noqa: F401
