"""Load the UNMODIFIED reference (intel/neural-compressor @ /root/reference) in-process.

TEST INFRASTRUCTURE ONLY.  Works only in the build container where
`/root/reference` is mounted; the GPU box has no reference tree, so nothing in
`tests -m gpu`, `smoke()` or `bench.py` imports this module.  It is used by
`oracle/gen_golden.py` (to produce the committed fixtures in `tests/golden/`)
and by the `not gpu` test that pins `oracle/woq_oracle.py` against the live
reference when the tree is present.

Two imports are missing from this image:
  * `prettytable`  -> tiny stub class (only used to print the op-stats table,
    neural_compressor/common/utils/utility.py:28)
  * `accelerate`   -> stub module injected into sys.modules AFTER transformers
    is imported (neural_compressor/torch/algorithms/layer_wise/utils.py:24)
"""
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("B200WOQ_REFERENCE_ROOT", "/root/reference")
VENDORED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")  # made by oracle/build_ref.py


def reference_available() -> bool:
    """The live tree (build container only)."""
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "neural_compressor"))


def vendored_available() -> bool:
    """The verbatim copy under oracle/_ref (travels to the GPU box; used by bench.py's reference / cpu_baseline legs)."""
    return os.path.isdir(os.path.join(VENDORED_ROOT, "neural_compressor", "torch"))


def _install_stubs():
    if "prettytable" not in sys.modules:
        pt = types.ModuleType("prettytable")

        class PrettyTable:  # minimal printable table
            def __init__(self, *a, **k):
                self.field_names = []
                self._rows = []

            def add_row(self, row):
                self._rows.append(row)

            def __str__(self):
                return "\n".join(str(r) for r in [self.field_names] + self._rows)

            get_string = __str__

        pt.PrettyTable = PrettyTable
        pt.__spec__ = importlib.machinery.ModuleSpec("prettytable", None)
        sys.modules["prettytable"] = pt
    import transformers  # noqa: F401  (must precede the accelerate stub)

    if "accelerate" not in sys.modules:
        acc = types.ModuleType("accelerate")
        accu = types.ModuleType("accelerate.utils")

        def set_module_tensor_to_device(*a, **k):
            raise NotImplementedError("accelerate stub (oracle/ref_loader.py)")

        accu.set_module_tensor_to_device = set_module_tensor_to_device
        acc.utils = accu
        acc.__version__ = "0.0.0"
        acc.__spec__ = importlib.machinery.ModuleSpec("accelerate", None)
        accu.__spec__ = importlib.machinery.ModuleSpec("accelerate.utils", None)
        sys.modules["accelerate"] = acc
        sys.modules["accelerate.utils"] = accu


def load_smooth_quant_utility():
    """The reference's SmoothQuant module (torch/algorithms/smooth_quant/utility.py) with `intel_extension_for_pytorch`
    and `peft` stubbed: the module hard-imports IPEX, but the smoothing transform itself (`TorchSmoothQuant.transform`:
    calibration, `cal_scale`, absorb grouping, `SQLinearWrapper` and its static activation qparams, the QDQ simulation
    helpers) is plain torch and runs on the CPU.  Only IPEX's int8 kernels have no counterpart here."""
    load_reference()

    def stub(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
        return sys.modules[name]

    ipex = stub("intel_extension_for_pytorch", __version__="2.5.0")
    ipex.quantization = stub("intel_extension_for_pytorch.quantization")
    stub("peft", PeftModel=type("PeftModel", (), {}))
    import neural_compressor.torch.algorithms.smooth_quant.utility as sq_utility

    return sq_utility


def load_reference():
    """Return the imported `neural_compressor` package of the reference (CPU forced)."""
    if reference_available():
        root = REFERENCE_ROOT
    elif vendored_available():
        root = VENDORED_ROOT
    else:
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT} nor vendored under {VENDORED_ROOT}")
    os.environ.setdefault("INC_TARGET_DEVICE", "cpu")  # torch/utils/auto_accelerator.py:436-442
    _install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    import neural_compressor  # noqa: F401
    import neural_compressor.torch.quantization  # noqa: F401

    return neural_compressor
