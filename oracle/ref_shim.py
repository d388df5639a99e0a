"""Container-only loader for the UNMODIFIED reference at /root/reference.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.  It exists so that
`oracle/make_golden.py` (run in the build container, where /root/reference is mounted read-only) can
execute the reference's own Python on CPU and write golden vectors under tests/golden/.  The GPU box
has no /root/reference; tests there read the committed fixtures instead.

The reference hard-codes CUDA in three places on this path; we patch *names at import time*, never files:
  * common.get_mask_from_lengths (common.py:86-97) allocates torch.cuda.LongTensor,
  * RADTTS.binarize_attention (radtts.py:320-334) passes device=attn.get_device() (== -1 on CPU),
  * RADTTS.infer (radtts.py:559,607,622,652) draws noise with torch.cuda.FloatTensor.
matplotlib is absent from the image and only imported for plotting (alignment.py:23) -> stub module.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("RADTTS_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "radtts.py"))


def load():
    """Returns a namespace with the reference modules (common, radtts, loss, alignment, splines, apm)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    import torch
    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if "matplotlib" not in sys.modules:
        m = types.ModuleType("matplotlib")
        m.use = lambda *a, **k: None
        m.pylab = types.ModuleType("matplotlib.pylab")
        m.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules["matplotlib"] = m
        sys.modules["matplotlib.pylab"] = m.pylab
        sys.modules["matplotlib.pyplot"] = m.pyplot
    import warnings
    warnings.filterwarnings("ignore")
    import common, radtts, loss, alignment, splines, attribute_prediction_model  # noqa: E401

    def mask_from_lengths(lengths):
        ids = torch.arange(int(lengths.max()), device=lengths.device)
        return ids < lengths.unsqueeze(1)

    for mod in (common, radtts, loss):
        mod.get_mask_from_lengths = mask_from_lengths
    try:
        import transformer
        transformer.get_mask_from_lengths = mask_from_lengths
    except Exception:
        pass

    def binarize_attention(self, attn, in_lens, out_lens):
        # same control flow as radtts.py:326-334, device-agnostic tensor creation
        with torch.no_grad():
            attn_cpu = attn.data.cpu().numpy()
            out = torch.zeros_like(attn)
            for b in range(attn.shape[0]):
                hard = radtts.mas(attn_cpu[b, 0, :int(out_lens[b]), :int(in_lens[b])])
                out[b, 0, :int(out_lens[b]), :int(in_lens[b])] = torch.tensor(hard, device=attn.device)
        return out

    radtts.RADTTS.binarize_attention = binarize_attention
    if not torch.cuda.is_available():
        torch.cuda.FloatTensor = torch.FloatTensor
    ns = types.SimpleNamespace(common=common, radtts=radtts, loss=loss, alignment=alignment,
                               splines=splines, apm=attribute_prediction_model, root=REF_ROOT)
    return ns
