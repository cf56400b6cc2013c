"""Import huggingface/transformers for the boundary tests: the reference checkout when present (authoring container),
otherwise whatever ``transformers`` is installed (the GPU box has no /root/reference)."""
import os
import sys
import types

REF_SRC = "/root/reference/src"


def import_transformers():
    if "transformers" not in sys.modules and os.path.isdir(REF_SRC) and not os.environ.get("B200_USE_INSTALLED_TRANSFORMERS"):
        sys.path.insert(0, REF_SRC)
        stub = types.ModuleType("transformers.dependency_versions_check")  # tokenizers version gate
        stub.dep_version_check = lambda *a, **k: None
        sys.modules.setdefault("transformers.dependency_versions_check", stub)
    import transformers

    return transformers
