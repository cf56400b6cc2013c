"""Locate huggingface/transformers for the boundary tests and both bench arms.

Order: the reference checkout (/root/reference/src, authoring container only) -> the UNMODIFIED reference installed by
`python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`
(git-ignored, travels to the GPU box with the snapshot) -> whatever `transformers` the image has.  The reference gates its
import on `tokenizers>=0.23.1` (src/transformers/dependency_versions_check.py:56); the image has 0.22.2 and the path
never tokenises, so that one module is pre-seeded with a stub.  Nothing else of the reference is touched."""
import os
import sys
import types

REF_SRC = "/root/reference/src"
REF_INSTALLED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def import_transformers():
    if "transformers" not in sys.modules and not os.environ.get("B200_USE_INSTALLED_TRANSFORMERS"):
        for root in (REF_SRC, REF_INSTALLED):
            if os.path.isdir(os.path.join(root, "transformers")):
                sys.path.insert(0, root)
                stub = types.ModuleType("transformers.dependency_versions_check")
                stub.dep_version_check = lambda *a, **k: None
                sys.modules.setdefault("transformers.dependency_versions_check", stub)
                break
    import transformers

    return transformers


def where():
    tf = import_transformers()
    return f"transformers {tf.__version__} ({os.path.dirname(tf.__file__)})"
