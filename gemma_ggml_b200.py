"""Import shim: the package directory is named `gemma.ggml_b200/` (after the reference repo), which is
not a valid Python identifier; `import gemma_ggml_b200` loads it from there."""
import importlib.util
import os
import sys

# GGB_PKG_DIR: load another build of the package (A/B runs of kernel variants on one GPU box, tools/ab.sh)
_dir = os.environ.get("GGB_PKG_DIR") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "gemma.ggml_b200")
_spec = importlib.util.spec_from_file_location("gemma_ggml_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gemma_ggml_b200"] = _mod
_spec.loader.exec_module(_mod)
