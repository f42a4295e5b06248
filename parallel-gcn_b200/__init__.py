"""parallel-gcn_b200: Blackwell-native (sm_100a) drop-in for the GCN training hot path of
davide-gurrieri/parallel-GCN.

Contents (only what the hot path needs):
  csrc/     hand-written CUDA kernels + the C ABI of include/gcnb.h  -> libgcn_b200.so
  host/     C++17 mirror of the reference's class API (Variable, Module chain, Adam, GCN, Parser) on top of the
            C ABI, plus the engine-level C entry points used by bench.py / tests
  binding.py  ctypes access to the C ABI (torch tensors are only device memory + stream plumbing)

The directory name carries a hyphen (as the task names it); import it with `__graft_entry__.load_package()`,
which registers it as `parallel_gcn_b200`.
"""
from . import build as _build  # noqa: F401  (build helper is importable without the library)


def __getattr__(name):
    if name == "binding":
        import importlib
        return importlib.import_module(__name__ + ".binding")
    raise AttributeError(name)
