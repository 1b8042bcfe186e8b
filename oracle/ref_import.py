"""Import the reference's own hot-path modules, unmodified, from /root/reference.

TEST INFRASTRUCTURE ONLY (see oracle/vatss_oracle.py header).  `import src.model` fails in
this image because src/model/__init__.py pulls voicefilter -> lipreader -> librosa (not
installed), so the package objects are synthesised and only the three hot-path files are
executed.  /root/reference is read-only: never write bytecode there.
/root/reference does not exist on the GPU box; callers must check `available()` first.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("VATSS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "model", "dptn_wav.py"))


def load():
    """Returns a namespace with the reference classes (model, loss)."""
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    sys.dont_write_bytecode = True
    if "src" not in sys.modules or not hasattr(sys.modules["src"], "__vatss_synth__"):
        src = types.ModuleType("src")
        src.__path__ = [os.path.join(REF_ROOT, "src")]
        src.__vatss_synth__ = True
        sys.modules["src"] = src
        pkg = types.ModuleType("src.model")
        pkg.__path__ = [os.path.join(REF_ROOT, "src", "model")]
        sys.modules["src.model"] = pkg
    dptn_wav = importlib.import_module("src.model.dptn_wav")
    dptn = importlib.import_module("src.model.dptn")
    dprnn = importlib.import_module("src.model.dprnn")
    losses = importlib.import_module("src.loss.ss_losses")
    ns = types.SimpleNamespace(
        DPTNAVWavEncDec=dptn_wav.DPTNAVWavEncDec,
        DPTNWavEncDec=dptn_wav.DPTNWavEncDec,
        DPTNEncDec=dptn.DPTNEncDec,
        DPRNNEncDec=dprnn.DPRNNEncDec,
        SplitToFolds=dprnn.SplitToFolds,
        OverlapAdd=dprnn.OverlapAdd,
        SiSNRLoss=losses.SiSNRLoss,
        SiSNRWavLoss=losses.SiSNRWavLoss,
    )
    return ns
