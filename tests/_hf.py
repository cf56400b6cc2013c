"""Import huggingface/transformers for the boundary tests (see baseline/ref_import.py for the search order: reference
checkout, then the unmodified reference installed under baseline/_ref — which is what the GPU box uses — then the image's)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline"))
from ref_import import REF_SRC, import_transformers, where  # noqa: E402,F401
