"""Import shims that let the *unmodified* reference (HenryHZY/VL-PET, /root/reference/src)
be imported under torch 2.11 / transformers 5.5 (the reference pins transformers 4.2.1).

Test infrastructure only.  Used by ``make_golden.py`` in the build
container, where /root/reference exists; never imported on the GPU box and never by the product.
"""
import os
import sys
import types

REF_SRC = os.environ.get("VLPET_REFERENCE_SRC", "/root/reference/src")


def available() -> bool:
    return os.path.isdir(REF_SRC)


def install_shims():
    import transformers
    import transformers.file_utils as fu
    import transformers.modeling_utils as mu

    fu.add_code_sample_docstrings = lambda *a, **k: (lambda f: f)

    def _stub(*a, **k):
        raise NotImplementedError("head pruning is not part of the oracle")

    mu.find_pruneable_heads_and_indices = _stub
    mu.prune_linear_layer = _stub
    if not hasattr(mu, "apply_chunking_to_forward"):
        try:
            from transformers.pytorch_utils import apply_chunking_to_forward
            mu.apply_chunking_to_forward = apply_chunking_to_forward
        except Exception:
            pass
    mp = types.ModuleType("transformers.utils.model_parallel_utils")
    mp.assert_device_map = lambda *a, **k: None
    mp.get_device_map = lambda *a, **k: None
    sys.modules["transformers.utils.model_parallel_utils"] = mp
    if not hasattr(mu.PreTrainedModel, "_vlpet_stock_init_weights"):
        mu.PreTrainedModel._vlpet_stock_init_weights = mu.PreTrainedModel.init_weights
    mu.PreTrainedModel.init_weights = lambda self: self.apply(self._init_weights)
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)


def remove_shims():
    """Give stock transformers models their own ``init_weights`` back (it ties the LM head to the embeddings; the shim above
    skips that).  The other shims only add names the installed transformers no longer has and stay."""
    import transformers.modeling_utils as mu
    stock = getattr(mu.PreTrainedModel, "_vlpet_stock_init_weights", None)
    if stock is not None:
        mu.PreTrainedModel.init_weights = stock


def import_bart_backbone():
    """-> the reference's forked ``my_transformers.modeling_bart`` module."""
    install_shims()
    import my_transformers.modeling_bart as mb
    return mb


def import_t5_backbone():
    install_shims()
    import my_transformers.modeling_t5 as mt
    return mt


def import_vl_bart():
    """-> the reference's ``src/modeling_bart.py`` (VisualEmbedding, VLBart...)."""
    mb = import_bart_backbone()
    import transformers
    import transformers.models.bart.modeling_bart as hfb
    hfb._make_causal_mask = mb._make_causal_mask
    hfb._expand_mask = mb._expand_mask

    class _Dummy:  # BeamScorer & co. are only used by generate()
        pass

    for n in ("BeamScorer", "BeamSearchScorer"):
        try:
            setattr(sys.modules["transformers"], n, _Dummy)
        except Exception:
            pass
    import modeling_bart as vlb
    return vlb


def import_adapters():
    install_shims()
    import adapters
    return adapters


def parse_args(flags):
    """Run the reference's own argparse (src/param.py:59-419) on a flag list."""
    install_shims()
    import param
    old = sys.argv
    sys.argv = ["x"] + list(flags)
    try:
        return param.parse_args()
    finally:
        sys.argv = old
