"""The shared synthetic-audio recipe (qwen3-asr-swift_b200/q3asr/synth.py: input data, not oracle arithmetic), re-exported so the
tests can keep writing `from oracle import synth`."""
import importlib.util
import os

_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qwen3-asr-swift_b200", "q3asr", "synth.py")
_spec = importlib.util.spec_from_file_location("_q3asr_synth", _path)
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
SEED = _mod.SEED
clip = _mod.clip
